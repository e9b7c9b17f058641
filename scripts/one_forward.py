"""Two plain (no CUDA graph) forwards of the conv net at 640x320x32 — the command ncu wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matryodshka_b200 import synth
from matryodshka_b200.runtime import NetEngine

H, W, P, ngf, B = 320, 640, 32, 64, int(os.environ.get("B", "1"))
wts = synth.net_weights(6 * P, 2 * P, ngf)
eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, "cuda", max_batch=B, precision=os.environ.get("PREC", "fp16_fp8x"))
hi, lo = eng.input_buffers(B)
hi.normal_(); lo.zero_()
out = torch.empty((B, H, W, 2 * P), device="cuda")
rgba = torch.empty((B, H, W, P, 4), device="cuda")
fused = eng.can_fuse_rgba and os.environ.get("FUSED", "1") != "0"   # the head as the pipeline runs it: fused RGBA assembly
for _ in range(int(os.environ.get("N", "2"))):
    if fused:
        eng.forward_rgba(hi_lo=(hi, lo), out=rgba)
    else:
        eng.forward(hi_lo=(hi, lo), out=out)
torch.cuda.synchronize()
print("ok", "fused head" if fused else "plain head", bool(torch.isnan(rgba if fused else out).any()))
