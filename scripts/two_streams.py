"""Experiment: N frame pipelines on N streams, graph replays interleaved -- do the small kernels of one
frame fill the barrier/tail bubbles of the other frame's conv kernels?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matryodshka_b200 import synth
from matryodshka_b200.runtime import MSIPipeline

H, W, P, ngf = 320, 640, 32, 64
K = int(os.environ.get("K", "100"))
wts = synth.net_weights(6 * P, 2 * P, ngf)
ref, src = synth.ods_pair(1, H, W)
tp = synth.target_positions(1)
for n in (1, 2, 3):
    pipes = [MSIPipeline(wts, H, W, P, ngf, batch=1, device="cuda") for _ in range(n)]
    streams = [torch.cuda.Stream() for _ in range(n)]
    for p, s in zip(pipes, streams):
        p.set_inputs(ref, src, tgt_pos=tp)
        torch.cuda.synchronize()
        with torch.cuda.stream(s):
            for _ in range(3):
                p.step()
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    for i in range(K):
        with torch.cuda.stream(streams[i % n]):
            pipes[i % n].step()
    for s in streams:
        main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"{n} pipeline(s) in flight: {ms:.4f} ms/frame, {1000.0 / ms:.1f} frames/s")
    del pipes
    torch.cuda.empty_cache()
