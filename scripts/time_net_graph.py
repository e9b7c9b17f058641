"""Times the conv net alone as a CUDA-graph replay (no CPU launch pacing), and LayerNorm-free /
stage-only variants selected by env vars understood by experimental builds."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from matryodshka_b200 import synth
from matryodshka_b200.runtime import NetEngine

H, W, P, ngf, B = 320, 640, 32, 64, int(os.environ.get("B", "1"))
prec = os.environ.get("PREC", "fp16x3")
wts = synth.net_weights(6 * P, 2 * P, ngf)
eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, "cuda", max_batch=B, precision=prec)
hi, lo = eng.input_buffers(B)
hi.normal_(); lo.zero_()
out = torch.empty((B, H, W, 2 * P), device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        eng.forward(hi_lo=(hi, lo), out=out)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    eng.forward(hi_lo=(hi, lo), out=out)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 30
e0.record()
for _ in range(n):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"net graph replay: {e0.elapsed_time(e1) / n:.4f} ms  (B={B}, {prec}, EXP={os.environ.get('MSI_EXP', '0')}, "
      f"nan={bool(torch.isnan(out).any())})")
