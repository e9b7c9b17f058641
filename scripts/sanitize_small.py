"""Small-shape run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from matryodshka_b200 import ops, synth
from matryodshka_b200.runtime import MSIPipeline, NetEngine
from matryodshka_b200.highres import high_res_rerender
from matryodshka_b200.msi import MSI

H, W, P, ngf = 16, 32, 32, 64
ref, src = synth.ods_pair(2, H, W)
wts = synth.net_weights(6 * P, 2 * P, ngf)
pipe = MSIPipeline(wts, H, W, P, ngf, batch=2, device="cuda", use_graph=False)
pipe.set_inputs(ref, src, tgt_pos=synth.target_positions(2))
pipe.step()
torch.cuda.synchronize()
simt = NetEngine(synth.net_weights(24, 8, 8), 16, 32, 24, 8, 8, "cuda", max_batch=1, conv_impl="simt")
simt.forward(torch.rand(1, 16, 32, 24, device="cuda"))
planes = MSI().inv_depths(1, 100, 4)
eye = synth.identity_poses(1)
bw = torch.rand(1, 8, 16, 4, device="cuda")
high_res_rerender(torch.rand(1, 24, 48, 3, device="cuda"), torch.rand(1, 24, 48, 3, device="cuda"), bw, bw, eye, eye,
                  synth.intrinsics(1), np.zeros((1, 3), np.float32), planes)
# row-(f) kernels: wrap-pad net (msi_train_net), the other colour schemes, ODS and perspective renders
wts_w = synth.net_weights(6 * P, 2 * P, ngf, coord=False)
wrap = NetEngine(wts_w, H, W, 6 * P, 2 * P, ngf, "cuda", max_batch=2, variant="wrap")
wrap.forward(torch.rand(2, H, W, 6 * P, device="cuda") * 2 - 1)
psv = torch.rand(1, H, W, 6 * P, device="cuda") * 2 - 1
for which in ("blend_bg", "blend_bg_psv", "alpha_only"):
    n_out = ops.color_pred_channels(which, P)
    ops.rgba_assemble_ex(torch.rand(1, H, W, n_out, device="cuda") * 2 - 1, psv, which, P, want_weights=True)
rgba = torch.rand(1, H, W, 4, 4, device="cuda")
ops.render_perspective(rgba, np.array([[0.01, 0.0, -0.02]], np.float32), planes, psp_height=9, psp_width=20)
ops.render_ods(rgba, eye, 1.0, [0.032], planes)
ops.sweep_coords(np.tile(np.eye(4, dtype=np.float32).reshape(1, 1, 16), (1, 2, 1)), [0.032], planes, 1, 8, 16, "cuda")
torch.cuda.synchronize()
print("sanitize_small: done", float(pipe.out["rgb"].abs().mean()))
