"""How much of K1 / K4 / K5 of one frame hides behind the conv net of another frame?  Two streams, graph
replays of 8 net forwards and of M launches of the other stage, alone and together.
    python scripts/exp_overlap.py        (results: profiles/r1_v9_overlap_experiment.json)
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matryodshka_b200 import synth, runtime

dev = torch.device("cuda:0")
w = synth.net_weights(6 * 32, 64, 64, 8964)
pa = runtime.MSIPipeline(w, use_graph=False)
pb = runtime.MSIPipeline(w, use_graph=False)
ref, src = synth.ods_pair(1, 320, 640, 8964)
for p in (pa, pb):
    p.set_inputs(ref, src)
    p._enqueue()
torch.cuda.synchronize()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def graph_of(fn, reps):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    return g


stA = dict(pa._stages())
stB = dict(pb._stages())
R = 8
res = {}
g_net = graph_of(stA["net"], R)
for name, M in (("psv_build", 20), ("rgba_assemble", 40), ("render_composite", 30)):
    g_o = graph_of(stB[name], M)

    def run(a, b):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        sa.wait_event(e0); sb.wait_event(e0)
        if a:
            with torch.cuda.stream(sa):
                g_net.replay()
        if b:
            with torch.cuda.stream(sb):
                g_o.replay()
        e1.record(sa); e2.record(sb)
        torch.cuda.synchronize()
        return max(e0.elapsed_time(e1), e0.elapsed_time(e2))
    for _ in range(2):
        run(True, True)
    t_net = min(run(True, False) for _ in range(3))
    t_o = min(run(False, True) for _ in range(3))
    t_both = min(run(True, True) for _ in range(3))
    res[name] = dict(net_ms=round(t_net, 3), other_ms=round(t_o, 3), both_ms=round(t_both, 3),
                     hidden_frac=round((t_net + t_o - t_both) / min(t_net, t_o), 3))
print(json.dumps(res))
