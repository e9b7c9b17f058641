"""Runs the reference's OWN inference code on seeded inputs and writes tests/golden/reference_run.npz.

    python -m oracle.refrun.run_reference            (from the repo root, in the container that has /root/reference)

TEST INFRASTRUCTURE ONLY.  The reference (brownvc/matryodshka, /root/reference) is TensorFlow-1.14 graph
code that cannot be installed here; its Python source, however, parses and runs under Python 3.  This
script imports the reference's files UNMODIFIED from /root/reference -- geometry/{spherical,projector,
sampling}.py, matryodshka/{msi,nets}.py -- over a NumPy stand-in for the TensorFlow ops they call
(oracle/refrun/tf114_numpy.py), drives them the way test.py:110-170 does, and stores what they return.
tests/test_reference_golden.py then holds the restatement in oracle/*.py to these vectors (bit for bit
for the geometry, 1e-5 for the conv net whose summation order is BLAS's), and tests/test_gpu_reference_golden.py
holds the CUDA path to them.  /root/reference is read here only; the fixture travels to the GPU box.

What this pins: every line of the reference's own algorithm on the path (signs, swaps, bracketing, index
order, channel order, argument order, layer wiring, scope names).  What stays restated: the TensorFlow /
slim / tensorflow_graphics op semantics listed in tf114_numpy.py.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFERENCE = os.environ.get("MSI_REFERENCE_ROOT", "/root/reference")
OUT = os.environ.get("MSI_REFRUN_OUT") or os.path.join(REPO, "tests", "golden", "reference_run.npz")

SEED = 8964
SMALL = dict(H=16, W=32, P=4, NGF=8)       # same inputs as tests/golden/make_golden.py
TC = dict(H=16, W=32, P=32, NGF=64)
GEOM = dict(H=32, W=64, P=8)
HIGHRES = dict(h=8, w=16, Hh=40, Wh=96, P=4)
FULL = dict(H=320, W=640, P=32)            # BASELINE.json configs[1]: digests only


def load_reference():
    """Imports the reference modules from where they lie; returns (tf shim, spherical, projector, sampling, MSI)."""
    if not os.path.isdir(REFERENCE):
        raise SystemExit("reference tree %s not present: the fixture can only be regenerated where it is" % REFERENCE)
    from oracle.refrun import tf114_numpy as tf
    tf.install()
    # Python-2 implicit relative imports of the reference: `import spherical` inside geometry/, `from nets import`
    # inside matryodshka/
    for p in (os.path.join(REFERENCE, "matryodshka"), os.path.join(REFERENCE, "geometry"), REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import geometry.projector as pj            # noqa: E402  (the reference's, not this repo's)
    import geometry.spherical as spherical     # noqa: E402
    import geometry.sampling as sampling       # noqa: E402
    from matryodshka.msi import MSI            # noqa: E402
    for m in (pj, spherical, sampling, sys.modules[MSI.__module__]):
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(REFERENCE)), m.__file__
    return tf, spherical, pj, sampling, MSI


def set_test_flags(tf, **kw):
    """The flags test.py runs the low-res inference with (test.py:39-83, loader.py:30-42)."""
    flags = dict(input_type="ODS", operation="train", coord_net=True, net_only=False, supervision="",
                 transform_inverse_reg=False, jitter=False, ngf=SMALL["NGF"], which_color_pred="blend_psv", gcn=False)
    flags.update(kw)
    tf.set_flags(**flags)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def f32(a):
    return np.asarray(a, np.float32)


def run_small(tf, MSI, out):
    """test.py:110-159 on a 16x32, 4-plane, ngf-8 frame: infer_msi -> render view / depth -> uint8."""
    from matryodshka_b200 import synth
    H, W, P, NGF = SMALL["H"], SMALL["W"], SMALL["P"], SMALL["NGF"]
    ref, src = synth.ods_pair(1, H, W, SEED)
    eye = synth.identity_poses(1)
    intr = synth.intrinsics(1)
    tgt_pos = np.array([[0.03, -0.02, 0.04]], np.float32)
    model = MSI()
    planes = model.inv_depths(1, 100, P)
    out["small/planes"] = np.array(planes, np.float64)
    out["small/ref"], out["small/src"], out["small/tgt_pos"] = ref, src, tgt_pos
    tf.set_named_tensor("ref_pose_inv:0", np.linalg.inv(eye))
    T = lambda a: tf.convert_to_tensor(a, tf.float32)  # noqa: E731

    def infer(which, weights, coord_net=True, jitter_inv=None, extra="blend_weights_alphas_psv"):
        set_test_flags(tf, which_color_pred=which, coord_net=coord_net, ngf=NGF,
                       transform_inverse_reg=jitter_inv is not None, jitter=jitter_inv is not None)
        if jitter_inv is not None:
            tf.set_named_tensor("jitter_pose_inv:0", jitter_inv)
        tf.set_variables(weights)
        pred, net_input = model.infer_msi(T(src), T(ref), None, None, T(eye), T(eye), T(intr), which, P, planes,
                                          extra, ngf=NGF)
        missing = set(weights) - set(tf.variables_read)
        assert not missing, "variables the reference never read: %s" % sorted(missing)
        return pred, net_input

    n_out = {"blend_psv": 2 * P, "blend_bg": 3 + 2 * P, "blend_bg_psv": 3 + 3 * P, "alpha_only": P}
    w_coord = synth.net_weights(6 * P, 2 * P, NGF, SEED)
    pred, net_input = infer("blend_psv", w_coord)
    out["small/psv"] = f32(net_input)
    out["small/blend_weights"] = f32(pred["blend_weights"])
    out["small/alphas"] = f32(pred["alphas"])
    out["small/rgba_layers"] = f32(pred["rgba_layers"])
    rgba = pred["rgba_layers"]
    eye_rt = tf.expand_dims(tf.eye(4), axis=0)
    view = model.msi_render_equirect_view(rgba, eye_rt, T(tgt_pos), planes, T(intr))
    depth = model.msi_render_equirect_depth(rgba, eye_rt, T(tgt_pos), planes, T(intr))
    out["small/view"], out["small/depth"] = f32(view), f32(depth)
    out["small/view_u8"] = np.asarray(model.deprocess_image(view))
    out["small/depth_u8"] = np.asarray(model.deprocess_depth_image(depth))
    out["small/view_single"] = f32(model.msi_render_equirect_view_single(rgba, eye_rt, T(tgt_pos), planes, T(intr)))
    # a rotated target pose and a large offset
    rot = np.eye(4, dtype=np.float32)
    c, s = np.float32(np.cos(0.3)), np.float32(np.sin(0.3))
    rot[0, 0], rot[0, 2], rot[2, 0], rot[2, 2] = c, s, -s, c
    rot[:3, 3] = [0.01, -0.02, 0.015]
    big = np.array([[0.3, 0.1, -0.2]], np.float32)
    out["small/rot_pose"], out["small/big_pos"] = rot[None], big
    out["small/view_rot"] = f32(model.msi_render_equirect_view(rgba, T(rot[None]), T(big), planes, T(intr)))
    # the other renderers (row f4)
    for order in (1, -1):
        out["small/ods_view_%+d" % order] = f32(model.msi_render_ods_view(rgba, order, T(rot[None]), T(tgt_pos), planes, T(intr)))
    for vw, (ph, pw) in ((3, (27, 48)), (0, (20, 24))):
        out["small/psp_view_%d" % vw] = f32(model.msi_render_perspective_view(rgba, eye_rt, T(tgt_pos), planes, T(intr),
                                                                              viewing_window=vw, psp_height=ph, psp_width=pw))
    # the other colour schemes (msi.py:166-273) and the non-coord net (nets.py:387-469)
    for which in ("blend_bg", "blend_bg_psv", "alpha_only"):
        w = synth.net_weights(6 * P, n_out[which], NGF, SEED)
        # (asking blend_bg for 'blend_weights' hits an unbound local in the reference itself, msi.py:283-284)
        pred, _ = infer(which, w, extra="alphas_psv")
        out["small/%s/rgba_layers" % which] = f32(pred["rgba_layers"])
    w_plain = synth.net_weights(6 * P, 2 * P, NGF, SEED, coord=False)
    pred, _ = infer("blend_psv", w_plain, coord_net=False)
    out["small/train_net/rgba_layers"] = f32(pred["rgba_layers"])
    out["small/train_net/alphas"] = f32(pred["alphas"])
    # jittered inference (msi.py:1118-1120, test.py:141-146): ref_pose_inv = ref_pose_inv . jitter_pose_inv
    jit = np.eye(4, dtype=np.float32)
    c, s = np.float32(np.cos(0.02)), np.float32(np.sin(0.02))
    jit[1, 1], jit[1, 2], jit[2, 1], jit[2, 2] = c, -s, s, c
    jit[:3, 3] = [0.004, 0.002, -0.003]
    out["small/jitter_pose_inv"] = jit[None]
    pred, net_input = infer("blend_psv", w_coord, jitter_inv=jit[None])
    out["small/jitter/psv"] = f32(net_input)
    set_test_flags(tf)


def run_tc(tf, MSI, out):
    """The shape the tensor-core net runs (ngf 64, 32 planes, 192 input channels) on a 16x32 frame: both nets."""
    from matryodshka_b200 import synth
    H, W, P, NGF = TC["H"], TC["W"], TC["P"], TC["NGF"]
    ref, src = synth.ods_pair(1, H, W, SEED + 1)
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    tgt_pos = np.array([[-0.04, 0.01, 0.02]], np.float32)
    out["tc/tgt_pos"] = tgt_pos
    model = MSI()
    planes = model.inv_depths(1, 100, P)
    tf.set_named_tensor("ref_pose_inv:0", np.linalg.inv(eye))
    T = lambda a: tf.convert_to_tensor(a, tf.float32)  # noqa: E731
    for tag, coord in (("coord", True), ("plain", False)):
        set_test_flags(tf, coord_net=coord, ngf=NGF)
        tf.set_variables(synth.net_weights(6 * P, 2 * P, NGF, SEED, coord=coord))
        pred, net_input = model.infer_msi(T(src), T(ref), None, None, T(eye), T(eye), T(intr), "blend_psv", P, planes,
                                          "alphas", ngf=NGF)
        rgba = pred["rgba_layers"]
        out["tc/%s/rgba_layers" % tag] = f32(rgba)
        eye_rt = tf.expand_dims(tf.eye(4), axis=0)
        out["tc/%s/view" % tag] = f32(model.msi_render_equirect_view(rgba, eye_rt, T(tgt_pos), planes, T(intr)))
        out["tc/%s/depth" % tag] = f32(model.msi_render_equirect_depth(rgba, eye_rt, T(tgt_pos), planes, T(intr)))
    set_test_flags(tf)


def run_highres(tf, MSI, out):
    """The plane-streamed high-res re-render, test.py:296-383: the graph part (:300-340) through the reference's own
    MSI methods, the per-plane feed loop and the host-side over-composite (:354-383) restated here as test.py has
    them (they are script code inside main())."""
    from matryodshka_b200 import synth
    h, w, Hh, Wh, P = HIGHRES["h"], HIGHRES["w"], HIGHRES["Hh"], HIGHRES["Wh"], HIGHRES["P"]
    rng = np.random.default_rng(SEED + 5)
    hres_ref, hres_src = synth.ods_pair(1, Hh, Wh, SEED + 5)
    blend_weights = rng.uniform(0, 1, (1, h, w, P)).astype(np.float32)   # blend_weights.npy / alphas.npy of the low-res pass
    alphas = rng.uniform(0, 1, (1, h, w, P)).astype(np.float32)
    tgt_pos = np.array([[0.02, -0.03, 0.01]], np.float32)
    eye, intr = synth.identity_poses(1), synth.intrinsics(1)
    out["highres/blend_weights"], out["highres/alphas"], out["highres/tgt_pos"] = blend_weights, alphas, tgt_pos
    set_test_flags(tf)
    tf.set_named_tensor("ref_pose_inv:0", np.linalg.inv(eye))
    T = lambda a: tf.convert_to_tensor(a, tf.float32)  # noqa: E731
    model = MSI()
    psv_planes = model.inv_depths(1, 100, P)
    hres_ref_image = model.preprocess_image(T(hres_ref))
    hres_src_image = model.preprocess_image(T(hres_src))
    hres_output, hres_depth = None, 0.
    for i in range(P):
        curr_psv_plane = tf.slice(tf.constant(psv_planes), [i], [1])
        hres_net_input = model.format_network_input(hres_ref_image, hres_src_image, T(eye), T(eye), curr_psv_plane, T(intr))
        uw = tf.image.resize(T(blend_weights[:, :, :, i:i + 1]), [Hh, Wh], align_corners=True,
                             method=tf.image.ResizeMethod.BILINEAR)
        ucurr_alpha = tf.image.resize(T(alphas[:, :, :, i:i + 1]), [Hh, Wh], align_corners=True,
                                      method=tf.image.ResizeMethod.BILINEAR)
        ufg_rgb = hres_net_input[:, :, :, 0:3]
        ubg_rgb = hres_net_input[:, :, :, 3:6]
        ucurr_rgb = uw * ufg_rgb + (1 - uw) * ubg_rgb
        urgba_layers = tf.reshape(tf.concat([ucurr_rgb, ucurr_alpha], axis=3), [1, Hh, Wh, 1, 4])
        single = model.msi_render_equirect_view_single(urgba_layers, tf.expand_dims(tf.eye(4), axis=0), T(tgt_pos),
                                                       curr_psv_plane, T(intr))
        cur = np.asarray(single)[0].astype(np.float32)
        rgb, alpha = cur[:, :, :, :3], cur[:, :, :, 3:]
        alpha_as_depth = np.tile(cur[:, :, :, 3:], (1, 1, 1, 3))
        if i == 0:
            hres_output = rgb
        else:
            hres_output = hres_output * (1. - alpha) + rgb * alpha
            hres_depth = (i / P) * alpha_as_depth + hres_depth * (1.0 - alpha_as_depth)
    out["highres/output"] = f32(hres_output[0])
    out["highres/depth"] = f32(hres_depth[0])


def sweep_uv(tf, spherical, pj, H, W, depths, pose, order, baseline):
    """The coordinate half of projector.sweep_one (projector.py:138-158) for one batch item."""
    S, T = spherical.lat_long_grid([H, W])
    P = len(depths)
    intr = np.zeros((1, 3, 3), np.float32)
    intr[0] = np.array([[baseline, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    intrinsic = tf.concat([tf.convert_to_tensor(intr), tf.zeros([1, 1, 3], tf.float32)], axis=1)
    intrinsic = tf.concat([intrinsic, tf.zeros([1, 4, 1], tf.float32)], axis=2)
    intrinsic_tiled = tf.tile(intrinsic, [P, 1, 1])
    pose_tiled = tf.tile(tf.convert_to_tensor(pose[None], tf.float32), [P, 1, 1])
    points = spherical.backproject_spherical(S, T, tf.convert_to_tensor(depths, tf.float32), intrinsic_tiled)
    points = pj.apply_pose(points, pose_tiled)
    return np.asarray(spherical.project_ods(points, order, pose_tiled, intrinsic_tiled, W, H))


def run_geometry(tf, spherical, pj, sampling, MSI, out):
    H, W, P = GEOM["H"], GEOM["W"], GEOM["P"]
    depths = MSI().inv_depths(1, 100, P)
    eye = np.eye(4, dtype=np.float32)
    gen = np.eye(4, dtype=np.float32)
    a = 0.05
    gen[:3, :3] = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]], np.float32)
    gen[:3, 3] = [0.01, -0.005, 0.02]
    out["geom/general_pose"] = gen
    for tag, pose in (("eye", eye), ("gen", gen)):
        for order in (1, -1):
            out["geom/ods_uv_%s_%+d" % (tag, order)] = sweep_uv(tf, spherical, pj, H, W, depths, pose, order, 0.032)
    radius = tf.convert_to_tensor(depths, tf.float32)
    for k, pos in enumerate(([0.0, 0.0, 0.0], [0.03, -0.02, 0.04], [0.3, 0.1, -0.2])):
        uv = spherical.intersect_sphere(tf.convert_to_tensor(gen if k == 2 else eye), tf.convert_to_tensor(pos, tf.float32),
                                        radius, P, 1, W, H)
        out["geom/sphere_uv_%d" % k] = f32(uv)
    # resample with coordinates that wrap on every side (sampling.py:135-197)
    rng = np.random.default_rng(SEED)
    img = rng.uniform(-1, 1, (2, 6, 8, 3)).astype(np.float32)
    pix = np.stack([rng.uniform(-9, 17, (2, 5, 7)), rng.uniform(-7, 13, (2, 5, 7))], -1).astype(np.float32)
    out["geom/resample_img"], out["geom/resample_pix"] = img, pix
    out["geom/resample_out"] = f32(sampling.bilinear_wrapper2(tf.convert_to_tensor(img), tf.convert_to_tensor(pix)))


def run_full_digests(tf, spherical, pj, MSI, meta):
    """BASELINE.json configs[1] (320x640, 32 spheres): too large to store, so the fixture keeps digests of the
    reference's sweep coordinates and of the PSV of the bench's synthetic pair."""
    from matryodshka_b200 import synth
    H, W, P = FULL["H"], FULL["W"], FULL["P"]
    depths = MSI().inv_depths(1, 100, P)
    eye = np.eye(4, dtype=np.float32)
    full = {}
    for order in (1, -1):
        uv = sweep_uv(tf, spherical, pj, H, W, depths, eye, order, 0.032)
        invalid = (uv[..., 0] == 1.0) & (uv[..., 1] == 1.0)
        full["ods_uv_%+d" % order] = dict(
            sha256=digest(uv), invalid=int(invalid.sum()),
            floor_u_sha256=digest(np.floor(uv[..., 0]).astype(np.int32)),
            floor_v_sha256=digest(np.floor(uv[..., 1]).astype(np.int32)),
            u_min=float(uv[..., 0].min()), u_max=float(uv[..., 0].max()),
            v_min=float(uv[..., 1].min()), v_max=float(uv[..., 1].max()))
    ref, src = synth.ods_pair(1, H, W, SEED)
    model = MSI()
    tf.set_named_tensor("ref_pose_inv:0", np.eye(4, dtype=np.float32)[None])
    set_test_flags(tf)
    T = lambda a: tf.convert_to_tensor(a, tf.float32)  # noqa: E731
    psv = np.asarray(model.format_network_input(model.preprocess_image(T(ref)), model.preprocess_image(T(src)),
                                                T(synth.identity_poses(1)), T(synth.identity_poses(1)), depths,
                                                T(synth.intrinsics(1))))
    full["psv"] = dict(sha256=digest(psv), shape=list(psv.shape), sum=float(psv.astype(np.float64).sum()),
                       abs_sum=float(np.abs(psv.astype(np.float64)).sum()))
    radius = tf.convert_to_tensor(depths, tf.float32)
    tp = synth.target_positions(1, SEED)[0]
    uv = np.asarray(spherical.intersect_sphere(tf.convert_to_tensor(eye), tf.convert_to_tensor(tp, tf.float32), radius, P, 1, W, H))
    full["sphere_uv"] = dict(sha256=digest(uv), tgt_pos=[float(v) for v in tp],
                             floor_u_sha256=digest(np.floor(uv[..., 0]).astype(np.int32)),
                             floor_v_sha256=digest(np.floor(uv[..., 1]).astype(np.int32)))
    meta["full"] = full


def reference_revision():
    head = os.path.join(REFERENCE, ".git", "HEAD")
    try:
        ref = open(head).read().strip()
        if ref.startswith("ref:"):
            ref = open(os.path.join(REFERENCE, ".git", ref.split()[1])).read().strip()
        return ref
    except OSError:
        return "unknown (no .git in the reference tree); DESIGN.md names 831c407"


def main():
    tf, spherical, pj, sampling, MSI = load_reference()
    out, meta = {}, {}
    run_small(tf, MSI, out)
    run_tc(tf, MSI, out)
    run_highres(tf, MSI, out)
    run_geometry(tf, spherical, pj, sampling, MSI, out)
    run_full_digests(tf, spherical, pj, MSI, meta)
    meta.update(seed=SEED, small=SMALL, tc=TC, geom=GEOM, highres=HIGHRES, reference=REFERENCE, reference_revision=reference_revision(),
                numpy=np.__version__,
                files=["geometry/spherical.py", "geometry/projector.py", "geometry/sampling.py", "matryodshka/msi.py",
                       "matryodshka/nets.py"],
                file_sha256={f: hashlib.sha256(open(os.path.join(REFERENCE, f), "rb").read()).hexdigest()
                             for f in ("geometry/spherical.py", "geometry/projector.py", "geometry/sampling.py",
                                       "matryodshka/msi.py", "matryodshka/nets.py")})
    out["meta_json"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), np.uint8)
    np.savez_compressed(OUT, **out)
    print("wrote %s (%d arrays, %.1f KB)" % (OUT, len(out), os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
