"""Oracle (test infrastructure): restatement of the MSI facade's inference half.

Follows /root/reference matryodshka/msi.py: inv_depths (:1196-1217),
preprocess_image / deprocess_image / deprocess_depth_image (:1163-1194),
format_network_input (:1094-1130), the ``blend_psv`` branch of infer_msi
(:130-147, :276-289), msi_render_equirect_view (:407-429) and
msi_render_equirect_depth (:384-405).

The reference only works at batch 1 on this path (test.py:89);
batches here mean B independent frames, each run through the B=1 path.

PINNED to the reference's own code run over a restated TF op layer (tests/golden/reference_run.npz) -- see oracle/__init__.py.
"""
from __future__ import annotations

import numpy as np
import torch

from . import geometry_np as g
from . import net_torch

F32 = np.float32


def inv_depths(start_depth, end_depth, num_depths):
    """msi.py:1196-1217 -- python floats, sorted far -> near."""
    inv_start_depth = 1.0 / start_depth
    inv_end_depth = 1.0 / end_depth
    depths = [start_depth, end_depth]
    for i in range(1, num_depths - 1):
        fraction = float(i) / float(num_depths - 1)
        inv_depth = inv_start_depth + (inv_end_depth - inv_start_depth) * fraction
        depths.append(1.0 / inv_depth)
    depths = sorted(depths)
    return depths[::-1]


def preprocess_image(image, dt=F32):
    """msi.py:1163-1171.  [TF-1.14 convert_image_dtype]: float input is passed
    through unchanged, uint8 is scaled by 1/255."""
    image = np.asarray(image)
    if image.dtype == np.uint8:
        image = image.astype(dt) * dt(1.0 / 255)
    return image.astype(dt) * dt(2) - dt(1)


def convert_to_uint8(image):
    """[TF-1.14 convert_image_dtype(float -> uint8, saturate=False)]:
    ``cast(x * (255 + 0.5))`` -- C-style truncating cast, no saturation.
    Out-of-range values wrap the way a float->int32->uint8 cast does."""
    scaled = np.asarray(image, dtype=F32) * F32(255.5)
    return scaled.astype(np.int32).astype(np.uint8)


def deprocess_image(image, dt=F32):
    """msi.py:1173-1181."""
    image = (np.asarray(image, dtype=dt) + dt(1.0)) / dt(2.0)
    return convert_to_uint8(image)


def deprocess_depth_image(image):
    """msi.py:1186-1194 (no (x+1)/2)."""
    return convert_to_uint8(image)


def format_network_input(ref_image, src_image, ref_pose, src_pose, planes, intrinsics,
                         ref_pose_inv=None, jitter_pose_inv=None, dt=F32, return_aux=False):
    """msi.py:1094-1130 for one frame (B=1).  ``ref_pose_inv`` is the hidden
    graph tensor "ref_pose_inv:0" (:1115); ``jitter_pose_inv`` the optional
    "jitter_pose_inv:0" (:1118-1120).  Eye 0 = ref (order +1), eye 1 = src
    (order -1) (:1127).  Returns [1, H, W, 6P], channel = eye*3P + p*3 + rgb."""
    ref_pose = np.asarray(ref_pose, dtype=dt)
    src_pose = np.asarray(src_pose, dtype=dt)
    if ref_pose_inv is None:
        ref_pose_inv = np.linalg.inv(ref_pose.astype(np.float64)).astype(dt)
    ref_pose_inv = np.asarray(ref_pose_inv, dtype=dt)
    if jitter_pose_inv is not None:
        ref_pose_inv = _matmul44(ref_pose_inv, np.asarray(jitter_pose_inv, dtype=dt), dt)
    poses = [ref_pose, src_pose]
    images = [ref_image, src_image]
    outs, auxs = [], []
    for i in range(2):
        curr_pose = _matmul44(poses[i], ref_pose_inv, dt)
        psv, aux = g.sweep_one(images[i], 1 if (i % 2) == 0 else -1, planes, curr_pose,
                               intrinsics, dt, return_aux=True)
        outs.append(psv)
        auxs.append(aux[0])
    net_input = np.concatenate(outs, axis=3)
    if return_aux:
        return net_input, auxs
    return net_input


def _matmul44(a, b, dt):
    """Batched [1,4,4] x [1,4,4] product, k-sum left to right in dt."""
    a = np.asarray(a, dtype=dt).reshape(-1, 4, 4)
    b = np.asarray(b, dtype=dt).reshape(-1, 4, 4)
    out = np.zeros_like(a)
    for i in range(4):
        for j in range(4):
            acc = a[:, i, 0] * b[:, 0, j]
            for k in range(1, 4):
                acc = acc + a[:, i, k] * b[:, k, j]
            out[:, i, j] = acc
    return out


def assemble_rgba(msi_pred, net_input, num_msi_planes, dt=F32):
    """msi.py:130-147 (``blend_psv``): pred [B,H,W,2L] in (-1,1), net_input
    [B,H,W,6P] -> rgba_layers [B,H,W,L,4], blend_weights, alphas [B,H,W,L]."""
    L = num_msi_planes
    msi_pred = np.asarray(msi_pred, dtype=dt)
    net_input = np.asarray(net_input, dtype=dt)
    blend_weights = (msi_pred[..., :L] + dt(1.0)) / dt(2.0)
    alphas = (msi_pred[..., L:2 * L] + dt(1.0)) / dt(2.0)
    layers = []
    for i in range(L):
        fg_rgb = net_input[..., i * 3:(1 + i) * 3]
        bg_rgb = net_input[..., (L + i) * 3:(L + 1 + i) * 3]
        curr_alpha = alphas[..., i][..., None]
        w = blend_weights[..., i][..., None]
        curr_rgb = w * fg_rgb + (dt(1) - w) * bg_rgb
        layers.append(np.concatenate([curr_rgb, curr_alpha], axis=3))
    rgba_layers = np.stack(layers, axis=3).astype(dt)
    return rgba_layers, blend_weights, alphas


def color_pred_channels(which_color_pred, L):
    """msi.py:107-116."""
    return {"blend_psv": 2 * L, "blend_bg": 2 * L + 3, "blend_bg_psv": 3 * L + 3, "alpha_only": L}[which_color_pred]


def assemble_rgba_ex(msi_pred, net_input, num_msi_planes, which_color_pred, dt=F32):
    """msi.py:117-268: RGBA layers for every `which_color_pred`.  Returns (rgba_layers,
    blend_weights | None, alphas, bg_blend_weights | None)."""
    L = num_msi_planes
    if which_color_pred == "blend_psv":
        r, bw, al = assemble_rgba(msi_pred, net_input, L, dt)
        return r, bw, al, None
    msi_pred = np.asarray(msi_pred, dtype=dt)
    net_input = np.asarray(net_input, dtype=dt)
    one, two = dt(1.0), dt(2.0)
    blend_weights = bg_blend_weights = None
    layers = []
    if which_color_pred == "blend_bg":          # msi.py:166-190
        blend_weights = (msi_pred[..., :L] + one) / two
        alphas = (msi_pred[..., L:2 * L] + one) / two
        bg_rgb = msi_pred[..., -3:]
        for i in range(L):
            fg_rgb = net_input[..., i * 3:(1 + i) * 3]
            w = blend_weights[..., i][..., None]
            curr_rgb = w * fg_rgb + (one - w) * bg_rgb
            layers.append(np.concatenate([curr_rgb, alphas[..., i][..., None]], axis=3))
    elif which_color_pred == "blend_bg_psv":    # msi.py:209-247
        blend_weights = (msi_pred[..., :L] + one) / two
        bg_blend_weights = (msi_pred[..., 2 * L:3 * L] + one) / two
        alphas = (msi_pred[..., L:2 * L] + one) / two
        pred_bg = msi_pred[..., -3:]
        for i in range(L):
            fg_rgb = net_input[..., i * 3:(1 + i) * 3]
            bg_rgb = net_input[..., (L + i) * 3:(L + 1 + i) * 3]
            w = blend_weights[..., i][..., None]
            curr_rgb = w * fg_rgb + (one - w) * bg_rgb
            bg_w = bg_blend_weights[..., i][..., None]
            curr_rgb = bg_w * curr_rgb + (one - bg_w) * pred_bg
            layers.append(np.concatenate([curr_rgb, alphas[..., i][..., None]], axis=3))
    elif which_color_pred == "alpha_only":      # msi.py:249-268
        alphas = (msi_pred[..., :L] + one) / two
        for i in range(L):
            layers.append(np.concatenate([net_input[..., i * 3:(1 + i) * 3], alphas[..., i][..., None]], axis=3))
    else:
        raise ValueError(which_color_pred)
    return np.stack(layers, axis=3).astype(dt), blend_weights, alphas, bg_blend_weights


def infer_msi(raw_src_image, raw_ref_image, ref_pose, src_pose, intrinsics, num_msi_planes,
              psv_planes, weights, extra_outputs="", ngf=64, coord_net=True, which_color_pred="blend_psv"):
    """msi.py:40-289, ``blend_psv`` / ODS / operation=='train' path, one frame at
    a time (argument order src, ref as in the reference)."""
    preds = {"rgba_layers": [], "blend_weights": [], "alphas": [], "psv": []}
    B = raw_src_image.shape[0]
    for b in range(B):
        src = preprocess_image(raw_src_image[b:b + 1])
        ref = preprocess_image(raw_ref_image[b:b + 1])
        net_input = format_network_input(ref, src, ref_pose[b:b + 1], src_pose[b:b + 1],
                                         psv_planes, intrinsics[b:b + 1])
        net = net_torch.msi_coord_train_net if coord_net else net_torch.msi_train_net
        with torch.no_grad():
            pred = net(torch.from_numpy(net_input), color_pred_channels(which_color_pred, num_msi_planes), weights,
                       ngf=ngf).numpy()
        rgba, bw, al, _ = assemble_rgba_ex(pred, net_input, num_msi_planes, which_color_pred)
        preds["rgba_layers"].append(rgba)
        preds["blend_weights"].append(bw)
        preds["alphas"].append(al)
        preds["psv"].append(net_input)
    out = {"rgba_layers": np.concatenate(preds["rgba_layers"], 0)}
    if "blend_weights" in extra_outputs and "blend" in which_color_pred:
        out["blend_weights"] = np.concatenate(preds["blend_weights"], 0)
    if "alpha" in extra_outputs:
        out["alphas"] = np.concatenate(preds["alphas"], 0)
    net_input = np.concatenate(preds["psv"], 0)
    if "psv" in extra_outputs:
        out["psv"] = net_input
    return out, net_input


def _project_layers(rgba_layers, tgt_pose_rt, tgt_pos, planes, dt=F32):
    rgba_layers = np.asarray(rgba_layers, dtype=dt)
    B = np.asarray(tgt_pose_rt).shape[0]
    depths = np.tile(np.asarray(planes, dtype=dt).reshape(len(planes), 1), (1, B))
    layers = np.transpose(rgba_layers, (3, 0, 1, 2, 4))  # [L, B, H, W, 4]
    return g.projective_forward_sphere(layers, None, np.asarray(tgt_pose_rt, dtype=dt),
                                       np.asarray(tgt_pos, dtype=dt), depths, dt)


def msi_render_equirect_view(rgba_layers, tgt_pose_rt, tgt_pos, planes, intrinsics=None, dt=F32):
    """msi.py:407-429 -> [B, H, W, 3] in [-1, 1]."""
    proj = _project_layers(rgba_layers, tgt_pose_rt, tgt_pos, planes, dt)
    return g.over_composite([proj[i] for i in range(len(planes))], dt)


def msi_render_equirect_depth(rgba_layers, tgt_pose_rt, tgt_pos, planes, intrinsics=None, dt=F32):
    """msi.py:384-405 -> [B, H, W, 3] in [0, 1)."""
    proj = _project_layers(rgba_layers, tgt_pose_rt, tgt_pos, planes, dt)
    return g.over_composite_depth([proj[i] for i in range(len(planes))], dt)


def msi_render_ods_view(rgba_layers, order, jitter_pose, tgt_pos, planes, intrinsics, dt=F32):
    """msi.py:502-525 -> projector.projective_forward_ods (projector.py:101-127): the MSI seen from
    one ODS eye, over-composited.  jitter_pose [B,4,4]; intrinsics [B,3,3] (only [0][0][0] is read)."""
    rgba_layers = np.asarray(rgba_layers, dtype=dt)
    B, H, W, L, _ = rgba_layers.shape
    depths = np.asarray(planes, dtype=dt)
    layers = np.transpose(rgba_layers, (3, 0, 1, 2, 4))  # [L, B, H, W, 4]
    coords = [g.intersect_ods(np.asarray(jitter_pose, dtype=dt)[i], None, order, intrinsics, depths, L, B, W, H, dt)
              for i in range(B)]
    coords = np.stack(coords, axis=0).transpose(1, 0, 2, 3, 4)
    proj = [g.resample(layers[l], coords[l], dt) for l in range(L)]
    return g.over_composite(proj, dt)


def msi_render_perspective_view(rgba_layers, tgt_pose_rt, tgt_pos, planes, intrinsics=None, viewing_window=3,
                                psp_height=270, psp_width=480, dt=F32):
    """msi.py:475-500 -> projector.projective_forward_sphere_to_perspective (projector.py:64-99): the
    caller's pose is replaced by the viewing-window rotation; [B, psp_height, psp_width, 3]."""
    rgba_layers = np.asarray(rgba_layers, dtype=dt)
    B, H, W, L, _ = rgba_layers.shape
    depths = np.asarray(planes, dtype=dt)
    layers = np.transpose(rgba_layers, (3, 0, 1, 2, 4))  # [L, B, H, W, 4]
    pose = g.viewing_window_pose(viewing_window, dt)
    tgt_pos = np.asarray(tgt_pos, dtype=dt).reshape(B, 3)
    coords = np.stack([g.intersect_perspective(pose, tgt_pos[i], depths, L, B, W, H, psp_width, psp_height, None, dt)
                       for i in range(B)], axis=0).transpose(1, 0, 2, 3, 4)   # [L, B, h, w, 2]
    proj = [g.resample(layers[l], coords[l], dt) for l in range(L)]
    return g.over_composite(proj, dt)


def msi_render_equirect_view_single(rgba_layers, tgt_pose_rt, tgt_pos, planes, intrinsics=None, dt=F32):
    """msi.py:431-452: reprojected layers without compositing [L, B, H, W, 4]."""
    return _project_layers(rgba_layers, tgt_pose_rt, tgt_pos, planes, dt)
