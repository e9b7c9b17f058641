"""ODS-eye renderer (msi_render_ods_view) and the jittered (--transform_inverse_reg) inference path."""
import numpy as np
import pytest
import torch

from oracle import msi_np
from matryodshka_b200 import synth
from matryodshka_b200.msi import MSI, MSIConfig
from tests.test_gpu_geometry import _smooth_layers

pytestmark = pytest.mark.gpu
F32 = np.float32
DEV = "cuda"
TOL = 1e-3


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _small_rotation(a, t):
    return np.array([[np.cos(a), 0, np.sin(a), t[0]], [0, 1, 0, t[1]], [-np.sin(a), 0, np.cos(a), t[2]],
                     [0, 0, 0, 1]], F32)


@pytest.mark.parametrize("order", [1, -1])
def test_render_ods_view_matches_oracle(order):
    B, H, W, L = 2, 32, 64, 8
    rgba = _smooth_layers(B, H, W, L, 21)
    planes = msi_np.inv_depths(1, 100, L)
    pose = np.stack([np.eye(4, dtype=F32), _small_rotation(0.02, (0.004, -0.003, 0.002))])
    want = msi_np.msi_render_ods_view(rgba, order, pose, None, planes, synth.intrinsics(B))
    got = MSI().msi_render_ods_view(_t(rgba), order, pose, None, planes, synth.intrinsics(B)).cpu().numpy()
    assert np.abs(got - want).max() < TOL


def test_ods_views_of_the_two_eyes_differ_by_parallax():
    """Sanity: left and right eye renders of the same MSI differ, and both equal the mirrored
    un-warped composite when the baseline is zero."""
    H, W, L = 32, 64, 6
    rgba = _smooth_layers(1, H, W, L, 5)
    planes = msi_np.inv_depths(1, 100, L)
    eye = np.eye(4, dtype=F32)[None]
    m = MSI()
    left = m.msi_render_ods_view(_t(rgba), 1, eye, None, planes, synth.intrinsics(1, 0.2))
    right = m.msi_render_ods_view(_t(rgba), -1, eye, None, planes, synth.intrinsics(1, 0.2))
    assert (left - right).abs().max().item() > 1e-2
    zero = m.msi_render_ods_view(_t(rgba), 1, eye, None, planes, synth.intrinsics(1, 0.0)).cpu().numpy()
    from oracle import geometry_np as g
    flat = g.over_composite([rgba[:, :, :, l] for l in range(L)])
    # zero baseline: every ray starts at the centre, so the view is the MSI itself (theta = -atan2(-sinS, cosS) = S)
    assert np.abs(zero - flat).max() < 2e-3


def test_jittered_inference_path():
    """test.py:141-146,160-165: infer with the sweep pose jittered (ref_pose_inv . jitter_pose_inv) and
    render with the jitter pose; both through the mirror API, against the oracle."""
    H, W, P, ngf = 16, 32, 4, 8
    ref, src = synth.ods_pair(1, H, W)
    wts = synth.net_weights(6 * P, 2 * P, ngf)
    planes = msi_np.inv_depths(1, 100, P)
    jitter = _small_rotation(0.015, (0.003, 0.002, -0.004))[None]
    jitter_inv = np.linalg.inv(jitter.astype(np.float64)).astype(F32)
    eye = synth.identity_poses(1)
    tp = np.array([[0.02, -0.01, 0.03]], F32)
    # oracle
    psv = msi_np.format_network_input(ref * 2 - 1, src * 2 - 1, eye, eye, planes, synth.intrinsics(1),
                                      ref_pose_inv=eye, jitter_pose_inv=jitter_inv)
    from oracle import net_torch
    with torch.no_grad():
        pred = net_torch.msi_coord_train_net(torch.from_numpy(psv), 2 * P, wts, ngf=ngf).numpy()
    rgba_o, _, _ = msi_np.assemble_rgba(pred, psv, P)
    want = msi_np.msi_render_equirect_view(rgba_o, jitter, tp, planes)
    # ours
    m = MSI(weights=wts, config=MSIConfig(conv_impl="simt", ngf=ngf, transform_inverse_reg=True, jitter=True))
    out, net_input = m.infer_msi(_t(src), _t(ref), None, None, eye, eye, synth.intrinsics(1), "blend_psv", P, planes,
                                 ngf=ngf, ref_pose_inv=eye, jitter_pose_inv=jitter_inv)
    assert np.abs(net_input.cpu().numpy() - psv).max() < TOL
    got = m.msi_render_equirect_view(out["rgba_layers"], jitter, tp, planes).cpu().numpy()
    assert np.abs(got - want).max() < TOL


@pytest.mark.parametrize("viewing_window,psp", [(3, (27, 48)), (1, (30, 40)), (0, (270, 480))])
def test_render_perspective_view_matches_oracle(viewing_window, psp):
    """MSI.msi_render_perspective_view (msi.py:475-500): pinhole window of psp_height x psp_width pixels
    out of the ERP layers (output size differs from the layer size), two frames with different offsets."""
    B, H, W, L = 2, 32, 64, 8
    rgba = _smooth_layers(B, H, W, L, 33)
    planes = msi_np.inv_depths(1, 100, L)
    tp = np.array([[0.02, -0.01, 0.03], [0.0, 0.0, 0.0]], F32)
    want = msi_np.msi_render_perspective_view(rgba, None, tp, planes, None, viewing_window, psp[0], psp[1])
    got = MSI().msi_render_perspective_view(_t(rgba), None, tp, planes, None, viewing_window, psp[0], psp[1])
    assert tuple(got.shape) == (B, psp[0], psp[1], 3)
    assert np.abs(got.cpu().numpy() - want).max() < TOL
    # the window looks somewhere: it is not constant, and two viewing windows differ
    assert float(got.std()) > 1e-3
