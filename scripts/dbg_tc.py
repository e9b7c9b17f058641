"""GPU debug: tcgen05 net vs oracle and vs the SIMT back end, layer by layer."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import msi_np, net_torch
from matryodshka_b200 import synth
from matryodshka_b200.runtime import NetEngine

H, W, P, ngf, B = (int(a) for a in (sys.argv[1:6] if len(sys.argv) > 5 else (32, 64, 32, 64, 1)))
prec = sys.argv[6] if len(sys.argv) > 6 else "fp16x3"
rng = np.random.default_rng(0)
x = rng.uniform(-1, 1, (B, H, W, 6 * P)).astype(np.float32)
wts = synth.net_weights(6 * P, 2 * P, ngf)
with torch.no_grad():
    want, feats = net_torch.msi_coord_train_net(torch.from_numpy(x), 2 * P, wts, ngf=ngf, return_feats=True)
simt = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, "cuda", max_batch=B, conv_impl="simt")
ps = simt.forward(torch.from_numpy(x).cuda())
torch.cuda.synchronize()
print("simt vs oracle pred", (ps.cpu() - want).abs().max().item(), flush=True)
tc = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, "cuda", max_batch=B, conv_impl="tcgen05", precision=prec)
pt = tc.forward(torch.from_numpy(x).cuda())
torch.cuda.synchronize()
print("tc vs oracle pred", (pt.cpu() - want).abs().max().item(), "tc vs simt", (pt - ps).abs().max().item(), flush=True)
for scope, f in feats.items():
    a = tc.read_activation(scope, B).cpu()
    r_tc, r_s = tc.read_raw(scope, B).cpu(), simt.read_raw(scope, B).cpu()
    print(f"{scope:10s} act err {float((a - f).abs().max()):.3e}  raw tc-vs-simt {float((r_tc - r_s).abs().max()):.3e}"
          f"  raw|max| {float(r_s.abs().max()):.3f}  nan {int(torch.isnan(r_tc).sum())}", flush=True)
