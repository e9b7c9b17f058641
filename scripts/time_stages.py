"""Stage times (K1 sweep, net, K5 render) of one pipeline at 640x320x32, each stage timed alone: `reps` back-to-back
launches between two CUDA events.  A/B builds of the library: MSI_B200_LIB=<path to .so>."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matryodshka_b200 import synth, _lib
from matryodshka_b200.runtime import MSIPipeline
H, W, P, ngf = 320, 640, int(os.environ.get("P", "32")), 64
wts = synth.net_weights(6 * P, 2 * P, ngf)
ref, src = synth.ods_pair(1, H, W)
pipe = MSIPipeline(wts, H, W, P, ngf, batch=1, device="cuda", use_graph=False)
pipe.set_inputs(ref, src, tgt_pos=synth.target_positions(1))
pipe.step()
torch.cuda.synchronize()
best = {}
for _ in range(3):
    for k, v in pipe.stage_times(reps=20).items():
        best[k] = min(best.get(k, 1e9), v)
print(os.path.basename(_lib.LIB_PATH), {k: round(v * 1e3, 1) for k, v in best.items()}, "us")
