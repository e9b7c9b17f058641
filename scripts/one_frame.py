"""Two plain (no CUDA graph) frames of the whole path at 640x320x32 -- the command ncu wraps for K1/K4/K5."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matryodshka_b200 import synth
from matryodshka_b200.runtime import MSIPipeline
H, W, P, ngf = 320, 640, 32, 64
wts = synth.net_weights(6 * P, 2 * P, ngf)
ref, src = synth.ods_pair(1, H, W)
pipe = MSIPipeline(wts, H, W, P, ngf, batch=1, device="cuda", use_graph=False)
pipe.set_inputs(ref, src, tgt_pos=synth.target_positions(1))
for _ in range(2):
    pipe.step()
torch.cuda.synchronize()
print("ok")
