"""Pins the CPU oracle (oracle/*.py) to the REFERENCE'S OWN CODE.

tests/golden/reference_run.npz holds what /root/reference's unmodified geometry/{spherical,projector,
sampling}.py and matryodshka/{msi,nets}.py returned when they were executed in the build container over a
NumPy stand-in for the TensorFlow-1.14 ops they call (oracle/refrun/run_reference.py, committed with the
fixture).  Geometry must agree BIT FOR BIT (both sides are float32 NumPy, one rounding per op: any
difference is a misreading of the reference); the conv net within 1e-5 (different summation order).
Nothing here reads /root/reference.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import geometry_np as g
from oracle import msi_np
from matryodshka_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
NET_TOL = 1e-5


@pytest.fixture(scope="module")
def ref():
    z = np.load(os.path.join(HERE, "golden", "reference_run.npz"))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(bytes(d.pop("meta_json")).decode())
    return d


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_inputs(ref):
    m = ref["meta"]["small"]
    planes = [float(v) for v in ref["small/planes"]]
    return m["H"], m["W"], m["P"], m["NGF"], planes, synth.identity_poses(1), synth.intrinsics(1)


def test_fixture_inputs_are_the_synthetic_ones(ref):
    H, W, P, NGF, planes, eye, intr = small_inputs(ref)
    r, s = synth.ods_pair(1, H, W, ref["meta"]["seed"])
    assert np.array_equal(r, ref["small/ref"]) and np.array_equal(s, ref["small/src"])
    assert planes == msi_np.inv_depths(1, 100, P)           # MSI.inv_depths, msi.py:1196-1217


def test_psv_bit_exact(ref):
    H, W, P, NGF, planes, eye, intr = small_inputs(ref)
    psv = msi_np.format_network_input(msi_np.preprocess_image(ref["small/ref"]), msi_np.preprocess_image(ref["small/src"]),
                                      eye, eye, planes, intr)
    assert np.array_equal(psv, ref["small/psv"])
    jit = msi_np.format_network_input(msi_np.preprocess_image(ref["small/ref"]), msi_np.preprocess_image(ref["small/src"]),
                                      eye, eye, planes, intr, jitter_pose_inv=ref["small/jitter_pose_inv"])
    assert np.array_equal(jit, ref["small/jitter/psv"])
    assert not np.array_equal(jit, psv)


@pytest.mark.parametrize("which", ["blend_psv", "blend_bg", "blend_bg_psv", "alpha_only"])
def test_infer_msi_colour_schemes(ref, which):
    H, W, P, NGF, planes, eye, intr = small_inputs(ref)
    w = synth.net_weights(6 * P, msi_np.color_pred_channels(which, P), NGF, ref["meta"]["seed"])
    out, _ = msi_np.infer_msi(ref["small/src"], ref["small/ref"], eye, eye, intr, P, planes, w,
                              "blend_weights_alphas_psv", ngf=NGF, which_color_pred=which)
    key = "small/rgba_layers" if which == "blend_psv" else "small/%s/rgba_layers" % which
    assert np.abs(out["rgba_layers"] - ref[key]).max() <= NET_TOL
    if which == "blend_psv":
        assert np.abs(out["blend_weights"] - ref["small/blend_weights"]).max() <= NET_TOL
        assert np.abs(out["alphas"] - ref["small/alphas"]).max() <= NET_TOL


def test_train_net_without_coord_channel(ref):
    H, W, P, NGF, planes, eye, intr = small_inputs(ref)
    w = synth.net_weights(6 * P, 2 * P, NGF, ref["meta"]["seed"], coord=False)
    out, _ = msi_np.infer_msi(ref["small/src"], ref["small/ref"], eye, eye, intr, P, planes, w, "alphas", ngf=NGF,
                              coord_net=False)
    assert np.abs(out["rgba_layers"] - ref["small/train_net/rgba_layers"]).max() <= NET_TOL
    assert np.abs(out["alphas"] - ref["small/train_net/alphas"]).max() <= NET_TOL


@pytest.mark.parametrize("tag,coord", [("coord", True), ("plain", False)])
def test_full_width_net(ref, tag, coord):
    """ngf 64, 32 planes (the channel counts of BASELINE.json's net) on a 16x32 frame, both nets."""
    m = ref["meta"]["tc"]
    H, W, P, NGF = m["H"], m["W"], m["P"], m["NGF"]
    r, s = synth.ods_pair(1, H, W, ref["meta"]["seed"] + 1)
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    planes = msi_np.inv_depths(1, 100, P)
    w = synth.net_weights(6 * P, 2 * P, NGF, ref["meta"]["seed"], coord=coord)
    out, _ = msi_np.infer_msi(s, r, eye, eye, intr, P, planes, w, "alphas", ngf=NGF, coord_net=coord)
    assert np.abs(out["rgba_layers"] - ref["tc/%s/rgba_layers" % tag]).max() <= NET_TOL
    rgba = ref["tc/%s/rgba_layers" % tag]
    assert np.array_equal(msi_np.msi_render_equirect_view(rgba, eye, ref["tc/tgt_pos"], planes), ref["tc/%s/view" % tag])
    assert np.array_equal(msi_np.msi_render_equirect_depth(rgba, eye, ref["tc/tgt_pos"], planes), ref["tc/%s/depth" % tag])


def test_renderers_bit_exact_on_the_reference_layers(ref):
    """Rendering starts from the reference's own RGBA layers, so every output must match bit for bit."""
    H, W, P, NGF, planes, eye, intr = small_inputs(ref)
    rgba, tp = ref["small/rgba_layers"], ref["small/tgt_pos"]
    view = msi_np.msi_render_equirect_view(rgba, eye, tp, planes)
    depth = msi_np.msi_render_equirect_depth(rgba, eye, tp, planes)
    assert np.array_equal(view, ref["small/view"])
    assert np.array_equal(depth, ref["small/depth"])
    assert np.array_equal(msi_np.deprocess_image(view), ref["small/view_u8"])
    assert np.array_equal(msi_np.deprocess_depth_image(depth), ref["small/depth_u8"])
    assert np.array_equal(msi_np.msi_render_equirect_view_single(rgba, eye, tp, planes), ref["small/view_single"])
    assert np.array_equal(msi_np.msi_render_equirect_view(rgba, ref["small/rot_pose"], ref["small/big_pos"], planes),
                          ref["small/view_rot"])
    for order in (1, -1):
        ods = msi_np.msi_render_ods_view(rgba, order, ref["small/rot_pose"], tp, planes, intr)
        assert np.array_equal(ods, ref["small/ods_view_%+d" % order])
    for vw, (ph, pw) in ((3, (27, 48)), (0, (20, 24))):
        psp = msi_np.msi_render_perspective_view(rgba, eye, tp, planes, viewing_window=vw, psp_height=ph, psp_width=pw)
        assert np.array_equal(psp, ref["small/psp_view_%d" % vw])


def _oracle_sweep_uv(H, W, depths, pose, order, baseline=0.032):
    S, T = g.lat_long_grid((H, W))
    pts = g.backproject_spherical(S, T, np.asarray(depths, np.float32))
    pts = g.apply_pose(pts, np.tile(np.asarray(pose, np.float32)[None], (len(depths), 1, 1)))
    intr = np.zeros((len(depths), 4, 4), np.float32)
    intr[:, 0, 0] = baseline
    return g.project_ods(pts, order, None, intr, W, H)


def test_sweep_and_sphere_coordinates_bit_exact(ref):
    m = ref["meta"]["geom"]
    H, W, P = m["H"], m["W"], m["P"]
    depths = msi_np.inv_depths(1, 100, P)
    for tag, pose in (("eye", np.eye(4, dtype=np.float32)), ("gen", ref["geom/general_pose"])):
        for order in (1, -1):
            uv = _oracle_sweep_uv(H, W, depths, pose, order)
            assert np.array_equal(uv, ref["geom/ods_uv_%s_%+d" % (tag, order)]), (tag, order)
    radius = np.asarray(depths, np.float32)
    for k, pos in enumerate(([0.0, 0.0, 0.0], [0.03, -0.02, 0.04], [0.3, 0.1, -0.2])):
        pose = ref["geom/general_pose"] if k == 2 else np.eye(4, dtype=np.float32)
        uv = g.intersect_sphere(pose, np.asarray(pos, np.float32), radius, P, 1, W, H)
        assert np.array_equal(uv, ref["geom/sphere_uv_%d" % k]), k
    out = g.resample(ref["geom/resample_img"], ref["geom/resample_pix"])
    assert np.array_equal(out, ref["geom/resample_out"])


def test_full_size_digests(ref):
    """BASELINE.json configs[1] (320x640, 32 spheres): the reference's sweep coordinates, validity count, PSV of
    the bench's synthetic pair and render coordinates, by SHA-256 of the float32 bytes."""
    full = ref["meta"]["full"]
    H, W, P = 320, 640, 32
    depths = msi_np.inv_depths(1, 100, P)
    eye = np.eye(4, dtype=np.float32)
    for order in (1, -1):
        want = full["ods_uv_%+d" % order]
        uv = _oracle_sweep_uv(H, W, depths, eye, order)
        assert int(((uv[..., 0] == 1.0) & (uv[..., 1] == 1.0)).sum()) == want["invalid"] == 68409
        assert sha(np.floor(uv[..., 0]).astype(np.int32)) == want["floor_u_sha256"]
        assert sha(np.floor(uv[..., 1]).astype(np.int32)) == want["floor_v_sha256"]
        assert sha(uv) == want["sha256"]
    r, s = synth.ods_pair(1, H, W, ref["meta"]["seed"])
    psv = msi_np.format_network_input(msi_np.preprocess_image(r), msi_np.preprocess_image(s), synth.identity_poses(1),
                                      synth.identity_poses(1), depths, synth.intrinsics(1))
    assert list(psv.shape) == full["psv"]["shape"]
    assert sha(psv) == full["psv"]["sha256"]
    tp = np.asarray(full["sphere_uv"]["tgt_pos"], np.float32)
    assert np.array_equal(tp, synth.target_positions(1, ref["meta"]["seed"])[0])
    uv = g.intersect_sphere(eye, tp, np.asarray(depths, np.float32), P, 1, W, H)
    assert sha(uv) == full["sphere_uv"]["sha256"]


def test_high_res_rerender_bit_exact(ref):
    """test.py:296-383 (plane-streamed high-res re-render): the reference's format_network_input on ONE tensor plane
    and msi_render_equirect_view_single, the align-corners resize of the stand-in, test.py's host-side composite."""
    from oracle import highres_np
    m = ref["meta"]["highres"]
    hres_ref, hres_src = synth.ods_pair(1, m["Hh"], m["Wh"], ref["meta"]["seed"] + 5)
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    planes = msi_np.inv_depths(1, 100, m["P"])
    rgb, dep = highres_np.high_res_rerender(hres_ref, hres_src, ref["highres/blend_weights"], ref["highres/alphas"], eye, eye,
                                            intr, ref["highres/tgt_pos"], planes)
    assert np.array_equal(rgb, ref["highres/output"])
    assert np.array_equal(dep, ref["highres/depth"])


def test_fixture_names_the_reference_files(ref):
    meta = ref["meta"]
    assert set(meta["file_sha256"]) == {"geometry/spherical.py", "geometry/projector.py", "geometry/sampling.py",
                                        "matryodshka/msi.py", "matryodshka/nets.py"}
    assert all(len(v) == 64 for v in meta["file_sha256"].values())
