"""Pipeline-event trace of CTA 0 of one halo-kernel layer (MSI_TC_TRACE=<scope>): prints per-slot waits."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from matryodshka_b200 import synth, _lib
from matryodshka_b200.runtime import NetEngine

scope = os.environ.setdefault("MSI_TC_TRACE", "conv2_1")
H, W, P, ngf, B = 320, 640, 32, 64, 1
wts = synth.net_weights(6 * P, 2 * P, ngf)
eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, "cuda", max_batch=B)
hi, lo = eng.input_buffers(B)
hi.normal_(); lo.zero_()
out = torch.empty((B, H, W, 2 * P), device="cuda")
for _ in range(3):
    eng.forward(hi_lo=(hi, lo), out=out)
torch.cuda.synchronize()
lib = _lib.load()
R, NR = 1024, 10
buf = np.zeros(R * NR, dtype=np.int64)
lib.msi_debug_conv_trace.restype = ctypes.c_int
lib.msi_debug_conv_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
n = lib.msi_debug_conv_trace(buf.ctypes.data, buf.size)
assert n == buf.size, n
t = buf.reshape(NR, R)
t0 = t[7, 0]
def rel(a): return [int(x - t0) for x in a if x != 0]
full, comm, wemp, wtma, aemp, afull = (rel(t[i]) for i in (0, 1, 2, 3, 4, 5))
epi = rel(t[6]); tempty = rel(t[8])
print(f"layer {scope}: kernel body {int(t[7,2]-t0)} clk; mma warp starts at {int(t[7,1]-t0)}; slots {len(full)}, chunks {len(afull)}")
print("unit accumulators free (mma) at", tempty)
print("epilogue [tfull seen, stores done] per unit:", [(epi[i], epi[i+1]) for i in range(0, len(epi) - 1, 2)])
print("A: producer saw empty at", aemp)
print("A: mma saw full at      ", afull)
print("slot: W-empty seen | TMA issued | MMA saw full | MMAs+commit issued | full-to-full delta | TMA latency (issue->full seen)")
for i in range(len(full)):
    d = full[i] - full[i - 1] if i else 0
    print(f"{i:3d}: {wemp[i]:7d} {wtma[i]:7d} {full[i]:7d} {comm[i]:7d}   d={d:5d}  lat={full[i]-wtma[i]:5d}  issue={comm[i]-full[i]:4d}")

# ---- launch-to-launch timeline (nanoseconds, %globaltimer) of back-to-back launches of the same layer
from matryodshka_b200.runtime import profile_net_layers
buf0 = buf.copy()
profile_net_layers(eng, (hi, lo), out, reps=1)
torch.cuda.synchronize()
n = lib.msi_debug_conv_trace(buf.ctypes.data, buf.size)
g = buf.reshape(NR, R)[9]
nl = int(g[0])
print(f"\n{nl} launches recorded; events per launch (us, relative to the entry of the first listed launch):")
print("launch: entry | setup done | pdl_wait done | first A full | last MMA issued | epilogue loop done (CTA0) | stats done (CTA0) | CTA0 exit | grid's last CTA finalised")
base = None
for i in range(max(0, nl - 6), min(nl, 60)):
    ev = g[8 + 16 * i: 8 + 16 * i + 9]
    if base is None:
        base = ev[0]
    print(f"{i:3d}: " + "  ".join(f"{(int(x) - int(base)) / 1000.0:8.2f}" if x else "    -   " for x in ev))
