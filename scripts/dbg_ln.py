"""LayerNorm kernel against torch on the conv1_1 raw output of a small net (debugging aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from matryodshka_b200 import synth
from matryodshka_b200.runtime import NetEngine
H, W, P, ngf, B = 32, 64, 32, 64, 1
wts = synth.net_weights(6 * P, 2 * P, ngf)
for prec in ("fp16x3", "fp16_fp8x"):
    eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, "cuda", max_batch=B, precision=prec)
    x = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (B, H, W, 6 * P)).astype(np.float32)).cuda()
    eng.forward(x)
    for scope in ("conv1_1", "conv1_2", "conv2_2"):
        raw = eng.read_raw(scope, B).double()
        act = eng.read_activation(scope, B).double()
        g = torch.from_numpy(np.asarray(wts[f"net/{scope}/LayerNorm/gamma"])).cuda().double()
        b = torch.from_numpy(np.asarray(wts[f"net/{scope}/LayerNorm/beta"])).cuda().double()
        mean = raw.mean(dim=(1, 2, 3), keepdim=True)
        var = raw.var(dim=(1, 2, 3), keepdim=True, unbiased=False)
        want = torch.relu((raw - mean) / torch.sqrt(var + 1e-12) * g + b)
        err = (act - want).abs()
        print(prec, scope, "max err", float(err.max()), "per-channel max (first 16)", [round(float(v), 4) for v in err.amax(dim=(0, 1, 2))[:16]],
              "act max", float(act.max()), "want max", float(want.max()))
