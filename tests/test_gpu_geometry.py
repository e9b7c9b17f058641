"""GPU parity tests of the geometry stages (K1 psv_build, K4 rgba_assemble, K5 render_composite)
against the CPU oracle, through the C ABI (ctypes -> libmsi_b200.so).

Bars (north_star): the project_ods validity mask and the integer sample-index grid are
bit-exact w.r.t. the float32 oracle (index mismatches allowed only at knife-edge coordinates
within 1e-3 px of an integer, counted and bounded); float outputs within 1e-3 max-abs.
"""
import os

import numpy as np
import pytest
import torch

from oracle import geometry_np as g
from oracle import msi_np
from matryodshka_b200 import ops, synth
from matryodshka_b200.msi import MSI

pytestmark = pytest.mark.gpu
F32 = np.float32
DEV = "cuda"
TOL = 1e-3  # north_star tolerance on float RGBA


def _t(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def _index_parity(c_dev, c_ref, size):
    """floor() of device vs oracle coordinates; returns (#mismatch, #mismatch away from a knife edge)."""
    f_dev = np.floor(c_dev).astype(np.int64)
    f_ref = np.floor(c_ref).astype(np.int64)
    bad = f_dev != f_ref
    dist = np.abs(c_ref - np.round(c_ref))
    return int(bad.sum()), int((bad & (dist > 1e-3)).sum())


@pytest.mark.parametrize("H,W,P", [(32, 64, 4), (320, 640, 32)])
def test_sweep_coords_mask_bit_exact_and_index_grid(H, W, P):
    d = msi_np.inv_depths(1, 100, P)
    poses = np.tile(np.eye(4, dtype=F32).reshape(1, 1, 16), (1, 2, 1))
    uv, valid = ops.sweep_coords(poses, [0.032], d, 1, H, W, DEV)
    uv, valid = uv.cpu().numpy(), valid.cpu().numpy().astype(bool)
    S, T = g.lat_long_grid((H, W))
    pts = g.backproject_spherical(S, T, F32(d))
    total_bad = 0
    for e, order in enumerate((1, -1)):
        ref, aux = g.project_ods(pts, order, None, synth.intrinsics(1), W, H, return_aux=True)
        assert np.array_equal(valid[0, e], aux["valid"]), "disc<0 mask must be bit-exact"
        assert np.all(uv[0, e][~aux["valid"]] == 1.0)
        ok = aux["valid"]
        du = np.abs(uv[0, e][ok] - ref[ok])
        assert du.max() < 2e-3, du.max()
        for k, size in ((0, W), (1, H)):
            n_bad, n_far = _index_parity(uv[0, e][..., k][ok], ref[..., k][ok], size)
            assert n_far == 0, (e, k, n_bad, n_far)
            total_bad += n_bad
    # knife-edge floor flips only (identity pose puts many v on near-integers: SURVEY 0.7(iii))
    assert total_bad <= (2e-3 if H >= 320 else 3e-2) * uv.size, total_bad
    if H == 320:
        assert int((~valid[0, 0]).sum()) == 68409


@pytest.mark.parametrize("H,W,P,kind", [(32, 64, 4, "band"), (320, 640, 32, "band"), (64, 128, 8, "noise")])
def test_psv_build_matches_oracle(H, W, P, kind):
    ref, src = synth.ods_pair(1, H, W, kind=kind)
    d = msi_np.inv_depths(1, 100, P)
    want = msi_np.format_network_input(msi_np.preprocess_image(ref), msi_np.preprocess_image(src),
                                       synth.identity_poses(1), synth.identity_poses(1), d, synth.intrinsics(1))
    poses = np.tile(np.eye(4, dtype=F32).reshape(1, 1, 16), (1, 2, 1))
    got = ops.psv_build(_t(ref), _t(src), poses, [0.032], d, preprocess=True).cpu().numpy()
    err = np.abs(got - want)
    # white noise has unit gradient per pixel: a 1e-4 px coordinate difference is visible there
    assert err.max() < (TOL if kind == "band" else 5e-3), err.max()
    assert err.mean() < 1e-5
    # mirror API: MSI.format_network_input on preprocessed images
    m = MSI()
    got2 = m.format_network_input(_t(ref * 2 - 1), _t(src * 2 - 1), synth.identity_poses(1), synth.identity_poses(1),
                                  d, synth.intrinsics(1)).cpu().numpy()
    assert np.array_equal(got2, got)


def test_psv_build_uint8_input_and_hi_lo_operand():
    H, W, P = 32, 64, 32
    ref, src = synth.ods_pair(2, H, W)
    ref8, src8 = (ref * 255).astype(np.uint8), (src * 255).astype(np.uint8)
    d = msi_np.inv_depths(1, 100, P)
    want = np.concatenate([
        msi_np.format_network_input(msi_np.preprocess_image(ref8[b:b + 1]), msi_np.preprocess_image(src8[b:b + 1]),
                                    synth.identity_poses(1), synth.identity_poses(1), d, synth.intrinsics(1))
        for b in range(2)])
    poses = np.tile(np.eye(4, dtype=F32).reshape(1, 1, 16), (2, 2, 1))
    hi = torch.empty((2, H, W, 6 * P), dtype=torch.float16, device=DEV)
    lo = torch.empty_like(hi)
    got = ops.psv_build(_t(ref8), _t(src8), poses, [0.032, 0.032], d, preprocess=True, hi_lo=(hi, lo)).cpu().numpy()
    assert np.abs(got - want).max() < TOL
    rec = (hi.float() + lo.float()).cpu().numpy() / 16.0
    assert np.abs(rec - got).max() < 2e-6  # hi/lo split keeps ~22 bits


def test_psv_build_general_pose():
    """Non-identity sweep pose (the --transform_inverse_reg path, msi.py:1118-1120)."""
    H, W, P = 32, 64, 4
    ref, src = synth.ods_pair(1, H, W)
    d = msi_np.inv_depths(1, 100, P)
    a = 0.02
    pose = np.array([[np.cos(a), 0, np.sin(a), 0.004], [0, 1, 0, -0.003], [-np.sin(a), 0, np.cos(a), 0.002],
                     [0, 0, 0, 1]], F32)[None]
    want = g.sweep_one(ref * 2 - 1, 1, d, pose, synth.intrinsics(1))
    from matryodshka_b200.geometry import projector as pj
    got = pj.ods_sphere_sweep(_t(ref * 2 - 1), 1, d, pose, synth.intrinsics(1)).cpu().numpy()
    assert np.abs(got - want).max() < TOL


def test_rgba_assemble_bit_exact():
    B, H, W, L = 2, 16, 32, 8
    rng = np.random.default_rng(1)
    pred = rng.uniform(-1, 1, (B, H, W, 2 * L)).astype(F32)
    psv = rng.uniform(-1, 1, (B, H, W, 6 * L)).astype(F32)
    want, bw, al = msi_np.assemble_rgba(pred, psv, L)
    rgba, gbw, gal = ops.rgba_assemble(_t(pred), _t(psv), want_weights=True)
    assert np.array_equal(rgba.cpu().numpy(), want)
    assert np.array_equal(gbw.cpu().numpy(), bw) and np.array_equal(gal.cpu().numpy(), al)


def _smooth_layers(B, H, W, L, seed=0):
    rng = np.random.default_rng(seed)
    base = synth.band_limited_images(B * L, H, W, seed).reshape(B, L, H, W, 3).transpose(0, 2, 3, 1, 4)
    alpha = synth.band_limited_images(B * L, H, W, seed + 5)[..., :1].reshape(B, L, H, W, 1).transpose(0, 2, 3, 1, 4)
    return np.ascontiguousarray(np.concatenate([base * 2 - 1, alpha], -1).astype(F32))


@pytest.mark.parametrize("H,W,L,tp", [
    (32, 64, 4, (0.03, -0.02, 0.04)),
    (32, 64, 64, (0.05, 0.01, -0.03)),     # L=64: two sweeps of the lane<->layer mapping
    (64, 128, 32, (0.3, 0.1, -0.2)),       # SURVEY 8d edge case: large offset
    (320, 640, 32, (0.031, -0.017, 0.044)),
])
def test_render_composite_matches_oracle(H, W, L, tp):
    rgba = _smooth_layers(1, H, W, L)
    d = msi_np.inv_depths(1, 100, L)
    eye = np.eye(4, dtype=F32)[None]
    tp = np.array([tp], F32)
    want = msi_np.msi_render_equirect_view(rgba, eye, tp, d)
    wdep = msi_np.msi_render_equirect_depth(rgba, eye, tp, d)
    res = ops.render_composite(_t(rgba), eye, tp, d)
    got, gdep = res["rgb"].cpu().numpy(), res["depth"].cpu().numpy()
    assert np.abs(got - want).max() < TOL, np.abs(got - want).max()
    assert np.abs(gdep - wdep).max() < TOL
    # uint8 epilogue: convert_image_dtype semantics; a float difference can move a truncation edge by 1
    d8 = np.abs(res["rgb_u8"].cpu().numpy().astype(int) - msi_np.deprocess_image(want).astype(int))
    assert d8.max() <= 1 and (d8 > 0).mean() < 1e-2
    d8 = np.abs(res["depth_u8"].cpu().numpy().astype(int) - msi_np.deprocess_depth_image(wdep).astype(int))
    assert d8.max() <= 1
    # index grid of the reprojection
    uv = ops.intersect_sphere_coords(eye, tp, d, 1, H, W, DEV).cpu().numpy()[0]
    ref_uv = g.intersect_sphere(eye[0], tp[0], F32(d), L, 1, W, H)
    assert np.abs(uv - ref_uv).max() < 2e-3
    for k in range(2):
        n_bad, n_far = _index_parity(uv[..., k], ref_uv[..., k], W if k == 0 else H)
        assert n_far == 0 and n_bad <= 2e-3 * uv[..., k].size
    # ... and of the fast chain the fused kernel actually samples at (no IEEE divide / sqrt, polynomial atan2).
    # Contract: within 1e-3 px of the oracle wherever the longitude is well conditioned, floor() flips only at
    # knife-edge coordinates.  Where the hit point lies within 1e-3 R of the vertical axis (a handful of samples
    # around the two poles of the source sphere) one ulp of x or z already moves u by more than that -- for any
    # evaluation order, the strict chain included (2e-3 above) -- and the bound there is 5e-3 px.
    uvf = ops.intersect_sphere_coords(eye, tp, d, 1, H, W, DEV, fast=True).cpu().numpy()[0]
    phi_src = ref_uv[..., 1] / F32(H - 1) * F32(np.pi - np.pi / H) - F32(np.pi / 2 - np.pi / (2 * H))
    well = np.cos(phi_src) > 1e-3
    duv = np.abs(uvf - ref_uv)
    assert duv[well].max() < 1e-3 and duv.max() < 5e-3, (duv[well].max(), duv.max())
    assert (~well).sum() <= max(4, 1e-4 * well.size)
    flips = 0
    for k in range(2):
        n_bad, n_far = _index_parity(uvf[..., k], ref_uv[..., k], W if k == 0 else H)
        assert n_far == 0 and n_bad <= 2e-3 * uvf[..., k].size
        flips += n_bad
    print(f"render fast chain {H}x{W}x{L}: max |duv| = {duv[well].max():.2e} px ({duv.max():.2e} incl. {int((~well).sum())} "
          f"ill-conditioned samples), floor flips = {flips} of {uvf.size}")


def test_render_batch_and_pose_rotation():
    B, H, W, L = 3, 32, 64, 8
    rgba = _smooth_layers(B, H, W, L, 3)
    d = msi_np.inv_depths(1, 100, L)
    a = 0.1
    rot = np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]], F32)
    pose = np.stack([np.eye(4, dtype=F32), rot, rot.T.copy()])
    tp = synth.target_positions(B)
    want = msi_np.msi_render_equirect_view(rgba, pose, tp, d)
    got = MSI().msi_render_equirect_view(_t(rgba), _t(pose), _t(tp), d).cpu().numpy()
    assert np.abs(got - want).max() < TOL


def test_zero_offset_render_is_mirror_at_full_size():
    """Size-independent property: at tgt_pos = 0 the render is the horizontal mirror of
    the un-warped over-composite; every u is an integer +- float noise, the worst case for floor()."""
    H, W, L = 320, 640, 32
    rgba = _smooth_layers(1, H, W, L, 7)
    d = msi_np.inv_depths(1, 100, L)
    res = ops.render_composite(_t(rgba), np.eye(4, dtype=F32)[None], np.zeros((1, 3), F32), d)
    layers = _t(np.ascontiguousarray(rgba.transpose(3, 0, 1, 2, 4)))
    flat = ops.over_composite(layers).cpu().numpy()
    assert np.abs(res["rgb"].cpu().numpy() - flat[:, :, ::-1]).max() < TOL
    fdep = ops.over_composite(layers, depth_mode=True).cpu().numpy()
    assert np.abs(res["depth"].cpu().numpy() - fdep[:, :, ::-1]).max() < TOL


def test_alpha_one_shows_nearest_layer_and_project_layers():
    H, W, L = 32, 64, 6
    rgba = _smooth_layers(1, H, W, L, 9)
    rgba[..., 3] = 1.0
    d = msi_np.inv_depths(1, 100, L)
    eye = np.eye(4, dtype=F32)[None]
    tp = np.array([[0.02, 0.01, -0.01]], F32)
    m = MSI()
    proj = m.msi_render_equirect_view_single(_t(rgba), eye, tp, d)
    want = msi_np.msi_render_equirect_view_single(rgba, eye, tp, d)
    assert np.abs(proj.cpu().numpy() - want).max() < TOL
    out = m.msi_render_equirect_view(_t(rgba), eye, tp, d).cpu().numpy()
    # (project_layers evaluates the strict coordinate chain, the fused render its fast chain: ~1e-5 px apart)
    assert np.abs(out - proj[L - 1, ..., :3].cpu().numpy()).max() < 1e-5
    from matryodshka_b200.geometry import projector as pj
    comp = pj.over_composite([proj[l] for l in range(L)]).cpu().numpy()
    assert np.abs(comp - out).max() < 1e-5
    # same recurrence order in both kernels: from the SAME projected layers the composite is bit-identical
    rgba2 = _smooth_layers(1, H, W, L, 9)
    proj2 = m.msi_render_equirect_view_single(_t(rgba2), eye, np.zeros((1, 3), F32), d)
    assert np.array_equal(pj.over_composite([proj2[l] for l in range(L)]).cpu().numpy(),
                          ops.over_composite(proj2.contiguous()).cpu().numpy())


def test_resample_wraps_like_the_oracle():
    rng = np.random.default_rng(4)
    img = rng.uniform(-1, 1, (2, 9, 11, 5)).astype(F32)
    coords = rng.uniform(-15, 25, (2, 7, 6, 2)).astype(F32)
    from matryodshka_b200.geometry import sampling
    got = sampling.bilinear_wrapper2(_t(img), _t(coords)).cpu().numpy()
    assert np.array_equal(got, g.resample(img, coords))


def test_invalid_arguments_raise():
    from matryodshka_b200._lib import MsiError
    with pytest.raises(MsiError):
        ops.render_composite(torch.zeros((1, 8, 8, 4, 4), device=DEV), np.eye(4)[None], np.zeros((1, 3)),
                             [1.0, 2.0, 3.0, 4.0], want_rgb=False, want_depth=False)


@pytest.mark.parametrize("same_pose", [True, False])
def test_psv_two_eye_form_is_bit_identical_to_single_kernel_form(same_pose):
    """The RGBX + two-eyes-per-thread form (scratch given) must reproduce the one-thread-per-
    (pixel, eye, plane) form bit for bit, with shared and with distinct eye poses."""
    B, H, W, P = 2, 32, 64, 32
    ref, src = synth.ods_pair(B, H, W)
    d = msi_np.inv_depths(1, 100, P)
    poses = np.tile(np.eye(4, dtype=F32).reshape(1, 1, 16), (B, 2, 1))
    if not same_pose:
        poses[:, 1, 3] = 0.004
        poses[1, 0, 7] = -0.002
    hi_a = torch.empty((B, H, W, 6 * P), dtype=torch.float16, device=DEV)
    lo_a, hi_b, lo_b = torch.empty_like(hi_a), torch.empty_like(hi_a), torch.empty_like(hi_a)
    a = ops.psv_build(_t(ref), _t(src), poses, [0.032, 0.05], d, hi_lo=(hi_a, lo_a), use_scratch=True)
    b = ops.psv_build(_t(ref), _t(src), poses, [0.032, 0.05], d, hi_lo=(hi_b, lo_b), use_scratch=False)
    assert torch.equal(a, b) and torch.equal(hi_a, hi_b) and torch.equal(lo_a, lo_b)
    # uint8 images through the pre-pass
    r8, s8 = _t((ref * 255).astype(np.uint8)), _t((src * 255).astype(np.uint8))
    assert torch.equal(ops.psv_build(r8, s8, poses, [0.032, 0.05], d, use_scratch=True),
                       ops.psv_build(r8, s8, poses, [0.032, 0.05], d, use_scratch=False))


def test_stage_level_spherical_functions():
    """geometry.spherical.{lat_long_grid, backproject_spherical, project_ods, project_spherical,
    theta_phi_to_pixels} and projector.apply_pose, composed as sweep_one composes them."""
    from matryodshka_b200.geometry import projector as pj, spherical as sp
    H, W, P = 16, 32, 3
    d = F32([40.0, 6.0, 1.2])
    S, T = sp.lat_long_grid((H, W), device=DEV)
    So, To = g.lat_long_grid((H, W))
    assert np.array_equal(S.cpu().numpy(), So) and np.array_equal(T.cpu().numpy(), To)
    x, y, z = sp.backproject_spherical(S, T, d)
    xo, yo, zo = g.backproject_spherical(So, To, d)
    assert np.abs(x.cpu().numpy() - xo).max() < 1e-5 * 40 and np.abs(y.cpu().numpy() - yo).max() < 1e-5 * 40
    a = 0.05
    pose = np.array([[np.cos(a), 0, np.sin(a), 0.01], [0, 1, 0, 0.02], [-np.sin(a), 0, np.cos(a), -0.01],
                     [0, 0, 0, 1]], F32)[None]
    # apply_pose and the projections on the ORACLE's points: same inputs -> exact / few-ulp outputs
    pts = tuple(_t(t) for t in (xo, yo, zo))
    px, py, pz = pj.apply_pose(pts, pose)
    qx, qy, qz = g.apply_pose((xo, yo, zo), np.broadcast_to(pose, (P, 4, 4)))
    assert np.array_equal(px.cpu().numpy(), qx) and np.array_equal(pz.cpu().numpy(), qz)
    uv = sp.project_ods((_t(qx), _t(qy), _t(qz)), -1, None, synth.intrinsics(1), W, H).cpu().numpy()
    uvo, aux = g.project_ods((qx, qy, qz), -1, None, synth.intrinsics(1), W, H, return_aux=True)
    assert np.array_equal(uv[~aux["valid"]], uvo[~aux["valid"]])
    assert np.abs(uv - uvo)[aux["valid"]].max() < 2e-3
    us = sp.project_spherical((_t(qx), _t(qy), _t(qz)), 1, None, None, W, H).cpu().numpy()
    assert np.abs(us - g.project_spherical((qx, qy, qz), 1, None, None, W, H)).max() < 2e-3
    th = np.linspace(-3, 3, 50).astype(F32)
    ph = np.linspace(-1.5, 1.5, 50).astype(F32)
    tp = sp.theta_phi_to_pixels(_t(th), _t(ph), W, H).cpu().numpy()
    assert np.array_equal(tp, g.theta_phi_to_pixels(th, ph, W, H))


# ---- K1 with cached coordinates (msi_sweep_table_build + msi_psv_gather) -------------------------------
@pytest.mark.parametrize("B,H,W,P,shared", [(1, 320, 640, 32, True), (2, 32, 64, 32, True), (2, 32, 64, 8, False),
                                            (1, 10, 12, 4, True), (1, 16, 32, 6, True), (3, 24, 40, 64, False)])
def test_psv_gather_equals_psv_build_bit_for_bit(B, H, W, P, shared):
    """The table is written by the same device functions as the per-frame chain, so the gathered PSV -- float32 and
    the fp16 hi/lo operand, disc < 0 samples included -- is the SAME BITS as msi_psv_build's.  Cases: the full bench
    size; a batch sharing one rig (table_frames = 1) and one with a rig per frame (general poses, two baselines);
    a ragged last block (120 pixels, 64 per block); P = 6 (generic kernel); P = 64 (config 3)."""
    ref, src = synth.ods_pair(B, H, W, seed=5)
    d = msi_np.inv_depths(1, 100, P)
    poses = np.tile(np.eye(4, dtype=F32).reshape(1, 1, 16), (B, 2, 1))
    base = np.full((B,), 0.032, F32)
    if not shared:
        for b in range(B):
            a = 0.01 * (b + 1)
            m = np.array([[np.cos(a), 0, np.sin(a), 0.004 * b], [0, 1, 0, -0.003], [-np.sin(a), 0, np.cos(a), 0.002],
                          [0, 0, 0, 1]], F32)
            poses[b, 1] = m.reshape(16)
            base[b] = 0.032 + 0.004 * b
    cs = -(-6 * P // 64) * 64
    mk = lambda: (torch.zeros((B, H, W, cs), dtype=torch.float16, device=DEV), torch.zeros((B, H, W, cs), dtype=torch.float16, device=DEV))  # noqa: E731
    hl_a, hl_b = mk(), mk()
    want = ops.psv_build(_t(ref), _t(src), poses, base, d, preprocess=True, hi_lo=hl_a, c_stride=cs)
    tbl = ops.sweep_table(poses, base, d, H, W, DEV)
    assert tbl.frames == (1 if shared else B)
    got = ops.psv_gather(_t(ref), _t(src), tbl, preprocess=True, hi_lo=hl_b, c_stride=cs)
    assert torch.equal(got, want)
    assert torch.equal(hl_a[0], hl_b[0]) and torch.equal(hl_a[1], hl_b[1])
    # the table itself == msi_sweep_coords (which the mask / index-grid tests above hold to the oracle)
    uv, valid = ops.sweep_coords(poses[:tbl.frames], base[:tbl.frames], d, tbl.frames, H, W, DEV)   # [F,2,P,H,W,2]
    t = tbl.table.view(tbl.frames, H, W, P, 2, 2).permute(0, 4, 3, 1, 2, 5)
    assert torch.equal(t, uv)
    if (B, H, W, P) == (1, 320, 640, 32):
        assert int((valid[0, 0] == 0).sum()) == 68409 and bool((uv[0, 0][valid[0, 0] == 0] == 1.0).all())
    # uint8 images and the cache_coords switch of the tensor-level API
    ref8, src8 = _t((ref * 255).astype(np.uint8)), _t((src * 255).astype(np.uint8))
    assert torch.equal(ops.psv_build(ref8, src8, poses, base, d, cache_coords=True), ops.psv_build(ref8, src8, poses, base, d))


def test_sweep_table_cache_keyed_by_rig_values():
    H, W, P = 16, 32, 4
    d = msi_np.inv_depths(1, 100, P)
    poses = np.tile(np.eye(4, dtype=F32).reshape(1, 1, 16), (1, 2, 1))
    a = ops.sweep_table(poses, [0.032], d, H, W, DEV)
    assert ops.sweep_table(poses.copy(), [0.032], list(d), H, W, DEV) is a            # same values -> same table
    assert ops.sweep_table(poses, [0.033], d, H, W, DEV) is not a                    # another baseline
    p2 = poses.copy()
    p2[0, 1, 3] = 1e-4
    assert ops.sweep_table(p2, [0.032], d, H, W, DEV) is not a                       # another pose
    assert ops.sweep_table(poses, [0.032], msi_np.inv_depths(1, 50, P), H, W, DEV) is not a
