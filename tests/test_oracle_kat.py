"""Known-answer and self-consistency tests that pin the CPU oracle.

The reference ships no tests or golden vectors, so the oracle
is pinned by analytic properties of the algorithm, by hand-written direct-loop
restatements of the TF-1.14 conv / deconv index rules, by its float64 twin, and
by the committed fixtures under tests/golden/.
"""
import os

import numpy as np
import pytest
import torch

from oracle import geometry_np as g
from oracle import msi_np, net_torch
from matryodshka_b200 import synth

F32 = np.float32
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_inv_depths_kat():
    d = msi_np.inv_depths(1, 100, 32)
    assert len(d) == 32 and d[0] == 100 and d[-1] == 1
    assert all(d[i] > d[i + 1] for i in range(31))
    # uniform in inverse depth: 1/d[1] = 1/100 + (1 - 1/100)/31
    assert abs(d[1] - 1.0 / (0.01 + 0.99 / 31)) < 1e-12
    assert abs(d[1] - 23.846153846153847) < 1e-9
    assert msi_np.inv_depths(1, 100, 2) == [100, 1]


def test_lat_long_grid_pixel_centres():
    S, T = g.lat_long_grid((8, 16))
    assert S.shape == (8, 16) and S.dtype == F32
    assert np.allclose(S[0, 0], -np.pi + np.pi / 16) and np.allclose(S[0, -1], np.pi - np.pi / 16, atol=1e-6)
    assert np.allclose(T[0, 0], -np.pi / 2 + np.pi / 16) and np.all(S[0] == S[5]) and np.all(T[:, 0] == T[:, 3])
    # spacing is one pixel
    assert np.allclose(np.diff(S[0]), 2 * np.pi / 16, atol=1e-6)


def test_theta_phi_to_pixels_maps_grid_to_indices():
    H, W = 8, 16
    S, T = g.lat_long_grid((H, W), np.float64)
    uv = g.theta_phi_to_pixels(S, T, W, H, np.float64)
    jj, ii = np.meshgrid(np.arange(W), np.arange(H))
    assert np.allclose(uv[..., 0], jj, atol=1e-9) and np.allclose(uv[..., 1], ii, atol=1e-9)


def test_resample_identity_and_wrap():
    rng = np.random.default_rng(0)
    img = rng.uniform(-1, 1, (2, 5, 7, 3)).astype(F32)
    jj, ii = np.meshgrid(np.arange(7), np.arange(5))
    px = np.stack([jj, ii], -1).astype(F32)[None].repeat(2, 0)
    assert np.array_equal(g.resample(img, px), img)
    # shift by a whole image: floor-mod wrap in x and y (sampling.py:162-165)
    assert np.array_equal(g.resample(img, px + F32([7, 0])), img)
    assert np.array_equal(g.resample(img, px - F32([0, 5])), img)
    # half-pixel shift in x: average of neighbours with wrap-around
    half = g.resample(img, px + F32([0.5, 0]))
    assert np.allclose(half, 0.5 * (img + np.roll(img, -1, axis=2)), atol=1e-6)
    # x = -0.5 wraps to column W-1
    neg = g.resample(img[:1], np.full((1, 1, 1, 2), -0.5, F32) * F32([1, 0]))
    assert np.allclose(neg[0, 0, 0], 0.5 * (img[0, 0, 6] + img[0, 0, 0]), atol=1e-6)


def test_constant_image_gives_constant_psv_including_invalid_pixels():
    H, W, P = 16, 32, 4
    col = F32([0.3, -0.2, 0.9])
    img = np.broadcast_to(col, (1, H, W, 3)).astype(F32)
    d = msi_np.inv_depths(1, 100, P)
    psv, aux = g.sweep_one(img, 1, d, synth.identity_poses(1), synth.intrinsics(1), return_aux=True)
    assert psv.shape == (1, H, W, 3 * P)
    assert np.allclose(psv.reshape(1, H, W, P, 3), col, atol=1e-6)
    assert (~aux[0]["valid"]).sum() > 0  # pole rows have disc < 0 and snap to (1, 1)
    bad = aux[0]["uv"][~aux[0]["valid"]]
    assert np.all(bad == 1.0)


def test_project_ods_mirror_and_order_symmetry():
    H, W, P = 32, 64, 3
    d = F32([50.0, 5.0, 1.5])
    S, T = g.lat_long_grid((H, W))
    pts = g.backproject_spherical(S, T, d)
    uvl, auxl = g.project_ods(pts, 1, None, synth.intrinsics(1), W, H, return_aux=True)
    uvr, auxr = g.project_ods(pts, -1, None, synth.intrinsics(1), W, H, return_aux=True)
    ok = auxl["valid"] & auxr["valid"]
    jj = np.arange(W)[None, None, :]
    # u is the horizontal mirror (W-1-j) plus/minus a parallax shift
    shift_l = (uvl[..., 0] - (W - 1 - jj))
    shift_r = (uvr[..., 0] - (W - 1 - jj))
    mid = slice(H // 4, 3 * H // 4)
    okm = ok[:, mid]
    assert np.abs(shift_l[:, mid][okm] + shift_r[:, mid][okm]).max() < 1e-2  # opposite shifts
    # nearer spheres shift more
    m = [np.abs(shift_l[p, H // 2]).mean() for p in range(P)]
    assert m[0] < m[1] < m[2]
    # v is (nearly) the row index away from the poles
    assert np.abs(uvl[:, H // 2, :, 1] - H // 2).max() < 0.05


def _layers(B, H, W, L, seed=0):
    rng = np.random.default_rng(seed)
    rgba = rng.uniform(-1, 1, (B, H, W, L, 4)).astype(F32)
    rgba[..., 3] = rng.uniform(0, 1, (B, H, W, L)).astype(F32)
    return rgba


def test_zero_offset_render_is_mirrored_composite():
    B, H, W, L = 1, 16, 32, 4
    rgba = _layers(B, H, W, L)
    d = msi_np.inv_depths(1, 100, L)
    eye = np.eye(4)[None]
    out64 = msi_np.msi_render_equirect_view(rgba, eye, np.zeros((B, 3)), d, dt=np.float64)
    ref = g.over_composite([rgba[:, :, :, l].astype(np.float64) for l in range(L)], np.float64)
    assert np.allclose(out64, ref[:, :, ::-1], atol=1e-9)
    out32 = msi_np.msi_render_equirect_view(rgba, eye, np.zeros((B, 3), F32), d)
    assert np.abs(out32 - ref[:, :, ::-1]).max() < 2e-3  # u is integer +- f32 noise; bilinear is continuous


def test_alpha_extremes():
    B, H, W, L = 1, 8, 16, 5
    rgba = _layers(B, H, W, L, 1)
    d = msi_np.inv_depths(1, 100, L)
    eye = np.eye(4)[None]
    zero = np.zeros((B, 3))
    r0 = rgba.copy(); r0[..., 3] = 0
    out = msi_np.msi_render_equirect_view(r0, eye, zero, d, dt=np.float64)
    assert np.allclose(out, rgba[:, :, ::-1, 0, :3], atol=1e-9)  # alpha of layer 0 is ignored (:257-259)
    r1 = rgba.copy(); r1[..., 3] = 1
    out = msi_np.msi_render_equirect_view(r1, eye, zero, d, dt=np.float64)
    assert np.allclose(out, rgba[:, :, ::-1, L - 1, :3], atol=1e-9)
    dep = msi_np.msi_render_equirect_depth(r1, eye, zero, d, dt=np.float64)
    assert np.allclose(dep, (L - 1) / L)
    dep0 = msi_np.msi_render_equirect_depth(r0, eye, zero, d, dt=np.float64)
    assert np.allclose(dep0, 0.0)


def test_over_composite_recurrence():
    rng = np.random.default_rng(3)
    ls = [rng.uniform(0, 1, (1, 2, 2, 4)).astype(F32) for _ in range(4)]
    out = g.over_composite(ls)
    exp = ls[0][..., :3]
    for i in range(1, 4):
        a = ls[i][..., 3:]
        exp = ls[i][..., :3] * a + exp * (1 - a)
    assert np.allclose(out, exp, atol=1e-7)


def test_convert_to_uint8_truncates_without_saturation():
    x = F32([0.0, 1.0, 0.999, 0.5, 0.0039])
    assert msi_np.convert_to_uint8(x).tolist() == [0, 255, 255, 127, 0]
    assert msi_np.deprocess_image(F32([-1.0, 1.0, 0.0])).tolist() == [0, 255, 127]
    # no saturation: 1.01 * 255.5 = 258.05 -> 258 -> wraps to 2
    assert msi_np.convert_to_uint8(F32([1.01])).tolist() == [2]


def test_same_pad_rule():
    assert net_torch.same_pad(320, 3, 1) == (1, 1)
    assert net_torch.same_pad(320, 3, 2) == (0, 1)   # 0 before / 1 after
    assert net_torch.same_pad(40, 3, 1, 2) == (2, 2)
    assert net_torch.same_pad(7, 3, 2) == (1, 1)


def _direct_conv(x, w, stride, rate):
    """Direct loops: out[oy,ox,co] = sum x[oy*s + kh*r - pt, ox*s + kw*r - pl, ci] * w[kh,kw,ci,co]."""
    B, H, W, C = x.shape
    k = w.shape[0]
    pt, _ = net_torch.same_pad(H, k, stride, rate)
    pl, _ = net_torch.same_pad(W, k, stride, rate)
    Ho, Wo = -(-H // stride), -(-W // stride)
    out = np.zeros((B, Ho, Wo, w.shape[3]), np.float64)
    for oy in range(Ho):
        for ox in range(Wo):
            for kh in range(k):
                for kw in range(k):
                    iy, ix = oy * stride + kh * rate - pt, ox * stride + kw * rate - pl
                    if 0 <= iy < H and 0 <= ix < W:
                        out[:, oy, ox] += x[:, iy, ix].astype(np.float64) @ w[kh, kw].astype(np.float64)
    return out


@pytest.mark.parametrize("stride,rate", [(1, 1), (2, 1), (1, 2)])
def test_conv_same_matches_direct_loops(stride, rate):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(1, 6, 8, 3)).astype(F32)
    w = rng.normal(size=(3, 3, 3, 4)).astype(F32)
    y = net_torch.conv2d_same(torch.from_numpy(x), torch.from_numpy(w), stride, rate).numpy()
    assert np.allclose(y, _direct_conv(x, w, stride, rate), atol=1e-5)


def test_deconv_same_matches_direct_scatter():
    """4x4 stride-2 SAME transposed conv: oy = 2*iy - 1 + kh, weights [kh,kw,Cout,Cin]."""
    rng = np.random.default_rng(6)
    x = rng.normal(size=(1, 3, 4, 5)).astype(F32)
    w = rng.normal(size=(4, 4, 2, 5)).astype(F32)
    y = net_torch.conv2d_transpose_same(torch.from_numpy(x), torch.from_numpy(w)).numpy()
    out = np.zeros((1, 6, 8, 2))
    for iy in range(3):
        for ix in range(4):
            for kh in range(4):
                for kw in range(4):
                    oy, ox = 2 * iy - 1 + kh, 2 * ix - 1 + kw
                    if 0 <= oy < 6 and 0 <= ox < 8:
                        out[0, oy, ox] += w[kh, kw].astype(np.float64) @ x[0, iy, ix].astype(np.float64)
    assert np.allclose(y, out, atol=1e-5)


def test_layer_norm_is_global_over_hwc():
    rng = np.random.default_rng(7)
    x = torch.from_numpy(rng.normal(3.0, 2.0, size=(2, 4, 5, 6)).astype(F32))
    y = net_torch.layer_norm_relu(x, torch.ones(6), torch.zeros(6), relu=False)
    for b in range(2):
        assert abs(y[b].mean().item()) < 1e-5 and abs(y[b].var(unbiased=False).item() - 1) < 1e-4
    gam = torch.arange(1, 7, dtype=torch.float32)
    y2 = net_torch.layer_norm_relu(x, gam, torch.full((6,), 0.5), relu=False)
    assert torch.allclose(y2, y * gam + 0.5, atol=1e-5)


def test_coord_channel_is_abs_sin_latitude():
    c = net_torch.sph_coord_rows(5)
    assert np.allclose(c, [1, np.sin(np.pi / 4), 0, np.sin(np.pi / 4), 1], atol=1e-7)
    x = torch.zeros(1, 5, 3, 2)
    y = net_torch.add_sph_coords(x)
    assert y.shape == (1, 5, 3, 3) and torch.allclose(y[0, :, 1, 2], torch.from_numpy(c))


def test_net_flops_match_survey_table():
    assert abs(net_torch.net_gflop(320, 640, 192, 64) - 302.4) < 0.05
    assert abs(net_torch.net_gflop(320, 640, 384, 128) - 349.4) < 0.05
    assert abs(net_torch.net_gflop(640, 1280, 192, 64) - 1209.6) < 0.2


def test_f32_oracle_tracks_f64_twin():
    H, W, P = 32, 64, 4
    ref, src = synth.ods_pair(1, H, W)
    d = msi_np.inv_depths(1, 100, P)
    a32, aux32 = g.sweep_one(ref * 2 - 1, 1, d, synth.identity_poses(1), synth.intrinsics(1), return_aux=True)
    a64, aux64 = g.sweep_one(ref.astype(np.float64) * 2 - 1, 1, d, synth.identity_poses(1), synth.intrinsics(1),
                             np.float64, return_aux=True)
    both = aux32[0]["valid"] & aux64[0]["valid"]
    duv = np.abs(aux32[0]["uv"].astype(np.float64) - aux64[0]["uv"])[both]
    assert duv.max() < 5e-3  # up to 2.5e-3 px near the poles
    rgba = _layers(1, H, W, P, 2)
    tp = np.array([[0.03, -0.02, 0.04]])
    o32 = msi_np.msi_render_equirect_view(rgba, np.eye(4)[None], tp.astype(F32), d)
    o64 = msi_np.msi_render_equirect_view(rgba, np.eye(4)[None], tp, d, dt=np.float64)
    assert np.abs(o32 - o64).mean() < 1e-4


def test_golden_fixture_reproduces():
    path = os.path.join(GOLDEN, "msi_small.npz")
    z = np.load(path)
    from tests.golden.make_golden import run_small
    out = run_small()
    for k in ("psv", "pred", "rgba_layers", "render", "depth"):
        assert np.abs(out[k] - z[k]).max() < 2e-5, k
    assert np.array_equal(out["render_u8"], z["render_u8"]) or \
        np.abs(out["render_u8"].astype(int) - z["render_u8"].astype(int)).max() <= 1


# ---- row-(f) restatements -------------------------------------------------------------------------
def test_train_net_encoder_is_equivariant_to_circular_shifts_in_x():
    """nets.msi_train_net pads every conv input circularly along the width (wrap_pad, nets.py:288-295)
    and normalises over the whole map, so rolling the panorama by 8 columns (three stride-2 levels)
    rolls every encoder activation (conv1_1 .. conv4_3) by 8 / 4 / 2 / 1 columns: an analytic property
    the zero-padded coord net does not have.  The decoder is only approximately equivariant: its
    deconvs normalise over the un-cropped (2H+10) x (2W+10) output (nets.py:431-436), whose 10 extra
    columns repeat a different part of the panorama after the roll."""
    P, ngf, H, W = 4, 8, 16, 32
    wts = synth.net_weights(6 * P, 2 * P, ngf, coord=False)
    x = torch.from_numpy(np.random.default_rng(3).uniform(-1, 1, (1, H, W, 6 * P)).astype(F32))
    with torch.no_grad():
        y, f = net_torch.msi_train_net(x, 2 * P, wts, ngf=ngf, return_feats=True)
        y8, f8 = net_torch.msi_train_net(torch.roll(x, 8, dims=2), 2 * P, wts, ngf=ngf, return_feats=True)
        wts_c = synth.net_weights(6 * P, 2 * P, ngf, coord=True)
        _, fc = net_torch.msi_coord_train_net(x, 2 * P, wts_c, ngf=ngf, return_feats=True)
        _, fc8 = net_torch.msi_coord_train_net(torch.roll(x, 8, dims=2), 2 * P, wts_c, ngf=ngf, return_feats=True)
    for scope, shift in [("conv1_1", 8), ("conv1_2", 4), ("conv2_2", 2), ("conv3_3", 1), ("conv4_3", 1)]:
        assert float((torch.roll(f[scope], shift, dims=2) - f8[scope]).abs().max()) < 1e-5, scope
    assert float((torch.roll(fc["conv4_3"], 1, dims=2) - fc8["conv4_3"]).abs().max()) > 1e-3   # zero padding breaks it
    d = float((torch.roll(y, 8, dims=2) - y8).abs().max())
    assert 0 < d < 0.2   # decoder: close, not exact (pre-crop LayerNorm)


def test_wrap_pad_kat():
    x = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(1, 2, 3, 4)[..., :1].repeat(1, 1, 1, 1)  # [1,2,3,1]
    p = net_torch.wrap_pad(x, 1, 1)[0, :, :, 0].numpy()
    row0, row1 = x[0, 0, :, 0].numpy(), x[0, 1, :, 0].numpy()
    assert p.shape == (4, 5)
    assert np.all(p[0] == 0) and np.all(p[-1] == 0)                       # zero rows above / below
    assert list(p[1]) == [row0[-1], *row0, row0[0]] and list(p[2]) == [row1[-1], *row1, row1[0]]  # circular columns


def test_color_scheme_known_answers():
    """msi.py:166-268: blend weight +1 (w = 1) shows the reference PSV, -1 (w = 0) the background;
    alpha_only copies the reference PSV; alphas are (pred + 1) / 2 in every scheme."""
    L, H, W = 3, 2, 4
    rng = np.random.default_rng(5)
    psv = rng.uniform(-1, 1, (1, H, W, 6 * L)).astype(F32)
    fg = psv[..., :3 * L].reshape(1, H, W, L, 3)
    bg = psv[..., 3 * L:].reshape(1, H, W, L, 3)
    a = rng.uniform(-1, 1, (1, H, W, L)).astype(F32)
    bgc = rng.uniform(-1, 1, (1, H, W, 3)).astype(F32)
    ones = np.ones((1, H, W, L), F32)
    # blend_bg: [w | alpha | bg rgb]
    r, bw, al, bgw = msi_np.assemble_rgba_ex(np.concatenate([ones, a, bgc], -1), psv, L, "blend_bg")
    assert np.array_equal(r[..., :3], fg) and np.array_equal(al, (a + 1) / 2) and bgw is None and np.all(bw == 1)
    r, *_ = msi_np.assemble_rgba_ex(np.concatenate([-ones, a, bgc], -1), psv, L, "blend_bg")
    assert np.array_equal(r[..., :3], np.broadcast_to(bgc[:, :, :, None, :], fg.shape))
    # blend_bg_psv: [w | alpha | bg_w | bg rgb]: bg_w = 1 reduces to blend_psv, bg_w = 0 to the background
    w = rng.uniform(-1, 1, (1, H, W, L)).astype(F32)
    r1, *_ = msi_np.assemble_rgba_ex(np.concatenate([w, a, ones, bgc], -1), psv, L, "blend_bg_psv")
    r0, _, _ = msi_np.assemble_rgba(np.concatenate([w, a], -1), psv, L)
    assert np.array_equal(r1, r0)
    r2, _, _, bgw = msi_np.assemble_rgba_ex(np.concatenate([w, a, -ones, bgc], -1), psv, L, "blend_bg_psv")
    assert np.array_equal(r2[..., :3], np.broadcast_to(bgc[:, :, :, None, :], fg.shape)) and np.all(bgw == 0)
    # alpha_only: [alpha]
    r, bw, al, _ = msi_np.assemble_rgba_ex(a, psv, L, "alpha_only")
    assert np.array_equal(r[..., :3], fg) and np.array_equal(r[..., 3], (a + 1) / 2) and bw is None
    assert msi_np.color_pred_channels("blend_bg_psv", 32) == 99 and msi_np.color_pred_channels("alpha_only", 32) == 32


def test_perspective_centre_ray_known_answer():
    """spherical.py:367-401: with no rotation and no offset the ray through the image centre is
    (0, 0, -0.05): it hits every sphere at (0, 0, -R), i.e. theta = -atan2(-R, 0) = +pi/2, phi = 0,
    whatever the radius: u = ((3 pi / 2 - pi / W) / (2 pi - 2 pi / W)) (W - 1), v = (H - 1) / 2."""
    W, H, tw, th = 64, 32, 9, 5   # odd target size: the centre pixel sits at S = T = 0
    uv = g.intersect_perspective(np.eye(4, dtype=F32), np.zeros(3, F32), np.array([1.0, 7.0, 100.0], F32), 3, 1, W, H,
                                 tw, th)
    assert uv.shape == (3, th, tw, 2)
    u_want = ((1.5 * np.pi - np.pi / W) / (2 * np.pi - 2 * np.pi / W)) * (W - 1)
    assert np.allclose(uv[:, th // 2, tw // 2, 0], u_want, atol=1e-3)
    assert np.allclose(uv[:, th // 2, tw // 2, 1], (H - 1) / 2.0, atol=1e-3)
    # left-right symmetric in v, and the viewing-window pose for window 0 is the identity
    assert np.allclose(uv[..., 1], uv[:, :, ::-1, 1], atol=1e-4)
    assert np.allclose(g.viewing_window_pose(0), np.eye(4), atol=1e-7)
    r3 = g.viewing_window_pose(3)[:3, :3]
    assert np.allclose(r3 @ r3.T, np.eye(3), atol=1e-6) and np.allclose(r3[0, 2], -1.0, atol=1e-6)
