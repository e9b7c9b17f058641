"""GPU parity against the REFERENCE'S OWN CODE: tests/golden/reference_run.npz holds what /root/reference's
unmodified Python returned when run over a NumPy stand-in for TensorFlow 1.14 (oracle/refrun/).  Everything
goes through the mirror API -> C ABI -> sm_100a kernels; tolerances are the north star's (mask bit-exact,
float outputs <= 1e-3 max-abs).  Nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest
import torch

from matryodshka_b200 import ops, synth
from matryodshka_b200.msi import MSI, MSIConfig

pytestmark = pytest.mark.gpu
F32 = np.float32
DEV = "cuda"
TOL = 1e-3


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


@pytest.fixture(scope="module")
def ref():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_run.npz"))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(bytes(d.pop("meta_json")).decode())
    return d


def _err(got, want):
    return float(np.abs(got.detach().cpu().numpy().astype(np.float64) - want).max())


def test_small_frame_end_to_end(ref):
    """test.py:127-159 on the 16x32 / 4-plane / ngf-8 frame (SIMT conv back end: ngf < 64)."""
    m = ref["meta"]["small"]
    H, W, P, NGF = m["H"], m["W"], m["P"], m["NGF"]
    planes = [float(v) for v in ref["small/planes"]]
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    msi = MSI(weights=synth.net_weights(6 * P, 2 * P, NGF, ref["meta"]["seed"]), config=MSIConfig(conv_impl="simt", ngf=NGF))
    assert planes == msi.inv_depths(1, 100, P)
    out, net_input = msi.infer_msi(_t(ref["small/src"]), _t(ref["small/ref"]), None, None, eye, eye, intr, "blend_psv", P,
                                   planes, "blend_weights_alphas_psv", ngf=NGF)
    assert _err(net_input, ref["small/psv"]) < TOL
    assert _err(out["blend_weights"], ref["small/blend_weights"]) < TOL
    assert _err(out["alphas"], ref["small/alphas"]) < TOL
    assert _err(out["rgba_layers"], ref["small/rgba_layers"]) < TOL
    res = msi.msi_render_equirect(out["rgba_layers"], np.eye(4, dtype=F32)[None], ref["small/tgt_pos"], planes)
    assert _err(res["rgb"], ref["small/view"]) < TOL
    assert _err(res["depth"], ref["small/depth"]) < TOL
    assert np.abs(res["rgb_u8"].cpu().numpy().astype(int) - ref["small/view_u8"].astype(int)).max() <= 1
    assert np.abs(res["depth_u8"].cpu().numpy().astype(int) - ref["small/depth_u8"].astype(int)).max() <= 1


def test_renderers_on_the_reference_layers(ref):
    P = ref["meta"]["small"]["P"]
    planes = [float(v) for v in ref["small/planes"]]
    rgba, tp = _t(ref["small/rgba_layers"]), ref["small/tgt_pos"]
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    msi = MSI()
    assert _err(msi.msi_render_equirect_view(rgba, eye, tp, planes), ref["small/view"]) < TOL
    assert _err(msi.msi_render_equirect_depth(rgba, eye, tp, planes), ref["small/depth"]) < TOL
    assert _err(msi.msi_render_equirect_view_single(rgba, eye, tp, planes), ref["small/view_single"]) < TOL
    assert _err(msi.msi_render_equirect_view(rgba, ref["small/rot_pose"], ref["small/big_pos"], planes), ref["small/view_rot"]) < TOL
    for order in (1, -1):
        got = msi.msi_render_ods_view(rgba, order, ref["small/rot_pose"], tp, planes, intr)
        assert _err(got, ref["small/ods_view_%+d" % order]) < TOL
    for vw, (ph, pw) in ((3, (27, 48)), (0, (20, 24))):
        got = msi.msi_render_perspective_view(rgba, eye, tp, planes, viewing_window=vw, psp_height=ph, psp_width=pw)
        assert _err(got, ref["small/psp_view_%d" % vw]) < TOL


@pytest.mark.parametrize("which", ["blend_bg", "blend_bg_psv", "alpha_only"])
def test_colour_schemes(ref, which):
    m = ref["meta"]["small"]
    H, W, P, NGF = m["H"], m["W"], m["P"], m["NGF"]
    planes = [float(v) for v in ref["small/planes"]]
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    wts = synth.net_weights(6 * P, ops.color_pred_channels(which, P), NGF, ref["meta"]["seed"])
    msi = MSI(weights=wts, config=MSIConfig(conv_impl="simt", ngf=NGF))
    out, _ = msi.infer_msi(_t(ref["small/src"]), _t(ref["small/ref"]), None, None, eye, eye, intr, which, P, planes,
                           "alphas", ngf=NGF)
    assert _err(out["rgba_layers"], ref["small/%s/rgba_layers" % which]) < TOL


def test_jittered_sweep(ref):
    m = ref["meta"]["small"]
    P = m["P"]
    planes = [float(v) for v in ref["small/planes"]]
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    msi = MSI()
    x = msi.format_network_input(msi.preprocess_image(_t(ref["small/ref"])), msi.preprocess_image(_t(ref["small/src"])),
                                 eye, eye, planes, intr, jitter_pose_inv=ref["small/jitter_pose_inv"])
    assert _err(x, ref["small/jitter/psv"]) < TOL


@pytest.mark.parametrize("tag,coord", [("coord", True), ("plain", False)])
def test_tensor_core_net_against_the_reference(ref, tag, coord):
    """ngf 64 / 32 planes on a 16x32 frame: the tcgen05 fp16x3 conv net (both nets of nets.py) -> RGBA layers ->
    rendered view and depth, against the reference's own output."""
    m = ref["meta"]["tc"]
    H, W, P, NGF = m["H"], m["W"], m["P"], m["NGF"]
    r, s = synth.ods_pair(1, H, W, ref["meta"]["seed"] + 1)
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    wts = synth.net_weights(6 * P, 2 * P, NGF, ref["meta"]["seed"], coord=coord)
    msi = MSI(weights=wts, config=MSIConfig(coord_net=coord), device=DEV)
    planes = msi.inv_depths(1, 100, P)
    out, _ = msi.infer_msi(_t(s), _t(r), None, None, eye, eye, intr, "blend_psv", P, planes, "alphas", ngf=NGF)
    assert _err(out["rgba_layers"], ref["tc/%s/rgba_layers" % tag]) < TOL
    res = msi.msi_render_equirect(out["rgba_layers"], np.eye(4, dtype=F32)[None], ref["tc/tgt_pos"], planes)
    assert _err(res["rgb"], ref["tc/%s/view" % tag]) < TOL
    assert _err(res["depth"], ref["tc/%s/depth" % tag]) < TOL


def test_sweep_and_sphere_coordinates(ref):
    """project_ods validity mask bit-exact; coordinates and the sample-index grid as in tests/test_gpu_geometry.py."""
    m = ref["meta"]["geom"]
    H, W, P = m["H"], m["W"], m["P"]
    depths = MSI().inv_depths(1, 100, P)
    for tag, pose in (("eye", np.eye(4, dtype=F32)), ("gen", ref["geom/general_pose"])):
        poses = np.stack([pose, pose])[None].reshape(1, 2, 16)
        uv, valid = ops.sweep_coords(poses, [0.032], depths, 1, H, W, DEV)
        uv, valid = uv.cpu().numpy(), valid.cpu().numpy().astype(bool)
        for e, order in enumerate((1, -1)):
            want = ref["geom/ods_uv_%s_%+d" % (tag, order)]
            want_valid = ~((want[..., 0] == 1.0) & (want[..., 1] == 1.0))
            assert np.array_equal(valid[0, e], want_valid), (tag, order)
            assert np.abs(uv[0, e] - want).max() < TOL
            off_edge = np.abs(want - np.round(want)) > 1e-3       # away from a knife edge the index grid is exact
            assert np.array_equal(np.floor(uv[0, e])[off_edge], np.floor(want)[off_edge])
    for k, pos in enumerate(([0.0, 0.0, 0.0], [0.03, -0.02, 0.04], [0.3, 0.1, -0.2])):
        pose = ref["geom/general_pose"] if k == 2 else np.eye(4, dtype=F32)
        got = ops.intersect_sphere_coords(pose[None], np.asarray(pos, F32)[None], depths, 1, H, W, DEV).cpu().numpy()
        assert np.abs(got[0] - ref["geom/sphere_uv_%d" % k]).max() < TOL
    out = ops.resample(_t(ref["geom/resample_img"]), _t(ref["geom/resample_pix"]))
    assert _err(out, ref["geom/resample_out"]) < 1e-5


def test_high_res_rerender(ref):
    """test.py:296-383 on the GPU (msi_highres_plane / msi_highres_composite) against the reference run."""
    from matryodshka_b200.highres import high_res_rerender
    m = ref["meta"]["highres"]
    hres_ref, hres_src = synth.ods_pair(1, m["Hh"], m["Wh"], ref["meta"]["seed"] + 5)
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    planes = MSI().inv_depths(1, 100, m["P"])
    rgb, dep = high_res_rerender(_t(hres_ref), _t(hres_src), _t(ref["highres/blend_weights"]), _t(ref["highres/alphas"]), eye,
                                 eye, intr, ref["highres/tgt_pos"], planes)
    assert _err(rgb, ref["highres/output"]) < TOL
    assert _err(dep, ref["highres/depth"]) < TOL
