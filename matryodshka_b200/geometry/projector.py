"""Mirror of geometry/projector.py for the functions on the MSI inference path."""
import numpy as np
import torch

from .. import ops
from .. import torch_ops  # noqa: F401  (registers torch.ops.msi.*)


def _baselines(intrinsics):
    k = intrinsics.detach().cpu().numpy() if torch.is_tensor(intrinsics) else np.asarray(intrinsics)
    return k.astype(np.float32).reshape(-1, 3, 3)[:, 0, 0]


def ods_sphere_sweep(image, order, depths, pose, intrinsics):
    """projector.py:209-211 (sweep_one :129-170 with the ODS functions).  image [B,H,W,3]
    (already preprocessed), order +1 / -1, pose [B,4,4] -> [B,H,W,3P], channel = p*3 + rgb."""
    B = image.shape[0]
    P = len(depths)
    dev = image.device
    # pose [B,4,4], or [1,4,4] broadcast over the batch (the reference sweeps frame i with psv_src_poses[i:i+1])
    pose = torch.as_tensor(pose, dtype=torch.float32).reshape(-1, 1, 16)
    if pose.shape[0] == 1 and B > 1:
        pose = pose.expand(B, 1, 16)
    pose = pose.repeat(1, 2, 1).contiguous().to(dev)
    base = _baselines(intrinsics)
    if base.shape[0] == 1 and B > 1:
        base = np.repeat(base, B)
    psv = torch.ops.msi.psv_build(image.contiguous(), image.contiguous(), pose, ops._dev_f32(base, dev),
                                  ops._dev_f32(list(depths), dev), False)
    e = 0 if order > 0 else 1
    return psv[..., e * 3 * P:(e + 1) * 3 * P].contiguous()


def sweep_one(image, order, depths, pose, intrinsics, st_fun=None, backproj_fun=None, proj_fun=None):
    """projector.py:129-170; only the ODS function triple is built (the callbacks are accepted
    for signature compatibility and must be the spherical ones or None)."""
    return ods_sphere_sweep(image, order, depths, pose, intrinsics)


def projective_forward_sphere(src_images, intrinsics, tgt_pose_rt, tgt_pos, depths):
    """projector.py:34-62.  src_images [L,B,H,W,4], depths [L,B] (columns identical) ->
    reprojected layers [L,B,H,W,4]."""
    rgba = src_images.permute(1, 2, 3, 0, 4).contiguous()
    d = depths[:, 0] if torch.is_tensor(depths) else np.asarray(depths)[:, 0]
    dev, B = rgba.device, rgba.shape[0]
    return torch.ops.msi.project_layers(rgba, ops._dev_f32(tgt_pose_rt, dev, (B, 16)), ops._dev_f32(tgt_pos, dev, (B, 3)),
                                        ops._dev_f32(d, dev, (-1,)))


def over_composite(rgbas):
    """projector.py:246-265.  list (back to front) of [B,H,W,4] -> [B,H,W,3]."""
    return torch.ops.msi.over_composite(torch.stack(list(rgbas), 0).contiguous(), False)


def over_composite_depth(rgbas):
    """projector.py:225-244."""
    return torch.ops.msi.over_composite(torch.stack(list(rgbas), 0).contiguous(), True)


def apply_pose(points, pose):
    """projector.py:275-291.  points = (x, y, z) each [P, H, W]; pose [P, 4, 4] (or [1, 4, 4])."""
    x, y, z = points
    P = x.shape[0]
    return tuple(t.reshape(x.shape) for t in ops.point_op(
        ops.OP_APPLY_POSE, x.reshape(P, -1), y.reshape(P, -1), z.reshape(P, -1), pose=pose))
