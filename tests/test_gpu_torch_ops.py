"""The C ABI through the PyTorch dispatcher (torch.ops.msi.*, matryodshka_b200/torch_ops.py) and under the reference's
module names (matryodshka.msi / geometry.*): same results as the tensor-level entry points of ops.py."""
import numpy as np
import pytest
import torch

from oracle import msi_np
from matryodshka_b200 import ops, synth, torch_ops
from matryodshka_b200.runtime import NetEngine

pytestmark = pytest.mark.gpu
DEV = "cuda"
F32 = np.float32


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def test_ops_through_the_dispatcher_equal_the_direct_calls():
    B, H, W, P, ngf = 1, 32, 64, 32, 64
    ref, src = synth.ods_pair(B, H, W, seed=41)
    d = msi_np.inv_depths(1, 100, P)
    poses = np.tile(np.eye(4, dtype=F32).reshape(1, 1, 16), (B, 2, 1))
    psv = torch.ops.msi.psv_build(_t(ref), _t(src), _t(poses), _t(F32([0.032])), _t(F32(d)), True)
    assert torch.equal(psv, ops.psv_build(_t(ref), _t(src), poses, [0.032], d))
    table = torch.ops.msi.sweep_table(_t(poses), _t(F32([0.032])), _t(F32(d)), H, W)
    assert torch.equal(torch.ops.msi.psv_gather(_t(ref), _t(src), table, True), psv)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    eng = NetEngine(wts, H, W, 6 * P, 2 * P, ngf, DEV)
    h = torch_ops.register_engine(eng)
    pred = torch.ops.msi.net_forward(psv, h)
    assert torch.equal(pred, eng.forward(psv))
    rgba, bw, al, bgw = torch.ops.msi.rgba_assemble(pred, psv, 0, P)
    want = ops.rgba_assemble(pred, psv, want_weights=True)
    assert torch.equal(rgba, want[0]) and torch.equal(bw, want[1]) and torch.equal(al, want[2]) and bgw.numel() == 0
    fused = torch.ops.msi.net_forward_rgba(psv, h)
    assert float((fused - rgba).abs().max()) < 1e-6      # (fused: assembled from the hi + lo operand, ~22 bits of the PSV)
    eye, tp = _t(np.eye(4, dtype=F32).reshape(1, 16)), _t(synth.target_positions(1, 41))
    rgb, depth, rgb8, dep8 = torch.ops.msi.render_composite(rgba, eye, tp, _t(F32(d)))
    res = ops.render_composite(rgba, eye, tp, d)
    assert torch.equal(rgb, res["rgb"]) and torch.equal(depth, res["depth"]) and torch.equal(rgb8, res["rgb_u8"])
    layers = torch.ops.msi.project_layers(rgba, eye, tp, _t(F32(d)))
    assert torch.equal(torch.ops.msi.over_composite(layers, False), ops.over_composite(layers))
    uv = torch.ops.msi.intersect_sphere_coords(eye, tp, _t(F32(d)), H, W, False)
    samp = torch.ops.msi.resample(rgba[:, :, :, 0, :].contiguous(), uv[:, 0].contiguous())
    assert torch.equal(samp, layers[0])
    torch_ops.release_engine(h)
    with pytest.raises(RuntimeError):
        torch.ops.msi.net_forward(psv, h)


def test_reference_module_names_run_the_path():
    """`from matryodshka.msi import MSI`, `import geometry.projector as pj` (reference test.py:27, msi.py:26-31)."""
    from matryodshka.msi import MSI
    from matryodshka import nets
    import geometry.projector as pj
    import geometry.sampling as sampling
    H, W, P, ngf = 32, 64, 32, 64
    ref, src = synth.ods_pair(1, H, W, seed=42)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    model = MSI().load_weights(wts)
    planes = model.inv_depths(1, 100, P)
    eye = synth.identity_poses(1)
    out, net_input = model.infer_msi(_t(src), _t(ref), None, None, eye, eye, synth.intrinsics(1), "blend_psv", P, planes,
                                     "blend_weights_alphas", ngf=ngf)
    assert set(out) == {"rgba_layers", "blend_weights", "alphas"}
    tp = synth.target_positions(1, 42)
    view = model.msi_render_equirect_view(out["rgba_layers"], np.eye(4, dtype=F32)[None], tp, planes)
    want, _ = msi_np.infer_msi(src, ref, eye, eye, synth.intrinsics(1), P, planes, wts, ngf=ngf)
    wview = msi_np.msi_render_equirect_view(want["rgba_layers"], np.eye(4, dtype=F32)[None], tp, planes)
    assert float(np.abs(view.cpu().numpy() - wview).max()) < 1e-3
    # stage-level API under the reference's names
    sweep = pj.ods_sphere_sweep(_t(ref * 2 - 1), 1, planes, eye, synth.intrinsics(1))
    assert torch.equal(sweep, net_input[..., :3 * P])
    pred = nets.msi_coord_train_net(net_input, 2 * P, ngf, weights=wts)
    assert pred.shape == (1, H, W, 2 * P)
    proj = pj.projective_forward_sphere(out["rgba_layers"].permute(3, 0, 1, 2, 4).contiguous(), None, eye, tp,
                                        np.asarray(planes, F32).reshape(P, 1))
    assert float((pj.over_composite([proj[l] for l in range(P)]) - view).abs().max()) < 1e-5
    c = torch.rand(1, 5, 7, 2, device=DEV) * 40 - 5
    assert sampling.bilinear_wrapper2(_t(ref), c).shape == (1, 5, 7, 3)
    # a [1,4,4] pose broadcasts over a batch (the reference sweeps frame i with psv_src_poses[i:i+1])
    ref2 = _t(np.concatenate([ref, src]) * 2 - 1)
    both = pj.ods_sphere_sweep(ref2, 1, planes, eye, synth.intrinsics(1))
    assert torch.equal(both[:1], sweep)
    # the coord_net choice is checked against the weights
    from matryodshka_b200.msi import MSIConfig
    from matryodshka_b200._lib import MsiError
    with pytest.raises(MsiError):
        MSI(weights=wts, config=MSIConfig(coord_net=False)).infer_msi(_t(src), _t(ref), None, None, eye, eye, synth.intrinsics(1),
                                                                      "blend_psv", P, planes, ngf=ngf)
