"""Mirror of geometry/spherical.py for the functions on the MSI inference path.

``lat_long_grid`` is host table logic (float32 NumPy, TF-1.14 LinSpace semantics) uploaded
to the GPU; the per-pixel projections run in csrc/geom_kernels.cu.
"""
import numpy as np
import torch

from .. import ops
from .. import torch_ops  # noqa: F401  (registers torch.ops.msi.*)


def lat_long_grid(shape, epsilon=1.0e-12, device="cuda"):
    """spherical.py:42-44 -> S, T [H, W] float32 (pixel-centre longitudes / latitudes)."""
    H, W = shape
    s, t = ops.lat_long_axes(H, W)
    S, T = np.meshgrid(s, t)
    return torch.from_numpy(S.copy()).to(device), torch.from_numpy(T.copy()).to(device)


def intersect_sphere(pos, center, radius, num_planes, num_batch, width, height, epsilon=1e-12):
    """spherical.py:268-326 -> uv [L, H, W, 2]: where the target-view ray of every ERP pixel
    hits each sphere, in source pixel coordinates.  pos [4,4], center [3], radius [L]."""
    dev = radius.device if torch.is_tensor(radius) else (pos.device if torch.is_tensor(pos) else "cuda")
    uv = torch.ops.msi.intersect_sphere_coords(ops._dev_f32(pos, dev, (1, 16)), ops._dev_f32(center, dev, (1, 3)),
                                               ops._dev_f32(radius, dev, (-1,)), height, width, False)
    return uv[0]


def project_ods_sweep(depths, pose, intrinsics, order, width, height, device="cuda"):
    """backproject_spherical (:116-129) + apply_pose + project_ods (:170-233) for the whole ERP
    grid: uv [P, H, W, 2] and the `disc >= 0` mask [P, H, W] for one eye (order = +1 / -1)."""
    pose = torch.as_tensor(pose, dtype=torch.float32).reshape(1, 1, 16).repeat(1, 2, 1)
    base = torch.as_tensor(intrinsics, dtype=torch.float32).reshape(-1, 3, 3)[:1, 0, 0]
    uv, valid = ops.sweep_coords(pose, base, depths, 1, height, width, device)
    e = 0 if order > 0 else 1
    return uv[0, e], valid[0, e].bool()


def theta_phi_to_pixels(theta, phi, width, height):
    """spherical.py:54-68 -> uv [..., 2] source pixel coordinates of the angles."""
    uv = ops.point_op(ops.OP_THETA_PHI_TO_PIXELS, theta.reshape(-1), phi.reshape(-1), H=height, W=width)
    return uv.reshape(tuple(theta.shape) + (2,))


def backproject_spherical(S, T, depth, intrinsics=None):
    """spherical.py:116-129.  S, T [H, W]; depth [P] -> (x, y, z) each [P, H, W]."""
    depth = torch.as_tensor(depth, dtype=torch.float32, device=S.device).reshape(-1)
    x, y, z = ops.point_op(ops.OP_BACKPROJECT_SPHERICAL, S.reshape(-1), T.reshape(-1), depth)
    shp = (depth.numel(),) + tuple(S.shape)
    return x.reshape(shp), y.reshape(shp), z.reshape(shp)


def project_ods(points, order, pose, intrinsics, width, height):
    """spherical.py:170-233, tuple branch: points = (x, y, z) each [P, H, W]; order +1 / -1;
    intrinsics[0][0][0] = ODS baseline.  Returns uv [P, H, W, 2] ((1, 1) where disc < 0)."""
    x, y, z = points
    k = torch.as_tensor(intrinsics, dtype=torch.float32).reshape(-1, 3, 3)
    uv = ops.point_op(ops.OP_PROJECT_ODS, x.reshape(-1), y.reshape(-1), z.reshape(-1), order=order,
                      baseline=float(k[0, 0, 0]), H=height, W=width)
    return uv.reshape(tuple(x.shape) + (2,))


def project_spherical(points, order, pose, intrinsics, width, height):
    """spherical.py:235-246."""
    x, y, z = points
    uv = ops.point_op(ops.OP_PROJECT_SPHERICAL, x.reshape(-1), y.reshape(-1), z.reshape(-1), H=height, W=width)
    return uv.reshape(tuple(x.shape) + (2,))
