"""High-res plane-streamed re-render -- the reference driver's ``--test_type high_res`` mode
(test.py:284-394) as one call.

For every PSV plane the reference runs a TF session that builds one 4096x2048 PSV plane for both
eyes, upsamples that plane's saved low-res blend weight / alpha, blends, reprojects the single layer
and returns it to the host, where NumPy over-composites (seconds per frame).  Here the same
per-plane streaming runs as two kernels per plane on the GPU (msi_highres_plane,
msi_highres_composite) with the running composite resident in HBM: never more than one high-res
RGBA layer (134 MB at 4096x2048) is alive.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr


def high_res_rerender(hres_ref_image, hres_src_image, blend_weights, alphas, ref_pose, src_pose, intrinsics, tgt_pos,
                      psv_planes, tgt_pose_rt=None, ref_pose_inv=None):
    """hres_*_image [1,Hh,Wh,3] float32 in [0,1] or uint8 (CUDA); blend_weights / alphas [1,h,w,P] (the
    low-res outputs of ``MSI.infer_msi``); tgt_pos [1,3]; psv_planes list[P].
    Returns (hres_output [Hh,Wh,3] float32 in [-1,1], hres_depth [Hh,Wh,3] float32 in [0,1))."""
    _lib.require_cuda()
    lib = _lib.load()
    from .msi import MSI
    dev = hres_ref_image.device
    _, Hh, Wh, _ = hres_ref_image.shape
    _, lh, lw, L = blend_weights.shape
    P = len(psv_planes)
    assert L == P and alphas.shape == blend_weights.shape
    if hres_ref_image.dtype == torch.uint8:
        dt = _lib.IMG_U8
    elif hres_ref_image.dtype == torch.float32:
        dt = _lib.IMG_F32
    else:
        raise _lib.MsiError(f"high_res_rerender: unsupported image dtype {hres_ref_image.dtype}")
    poses = ops._dev_f32(MSI()._sweep_poses(ref_pose, src_pose, ref_pose_inv), dev, (2, 16))
    k = intrinsics.detach().cpu().numpy() if torch.is_tensor(intrinsics) else np.asarray(intrinsics)
    baseline = ops._dev_f32(k.astype(np.float32).reshape(-1, 3, 3)[:1, 0, 0], dev, (1,))
    depths = ops._dev_f32(list(psv_planes), dev, (-1,))
    pose_rt = ops._dev_f32(tgt_pose_rt if tgt_pose_rt is not None else np.eye(4, dtype=np.float32), dev, (16,))
    pos = ops._dev_f32(tgt_pos, dev, (3,))
    tb = ops.erp_tables(Hh, Wh, dev)
    bw = blend_weights.contiguous().float()
    al = alphas.contiguous().float()
    ref = hres_ref_image.contiguous()
    src = hres_src_image.contiguous()
    layer = torch.empty((Hh, Wh, 4), dtype=torch.float32, device=dev)
    acc_rgb = torch.empty((Hh, Wh, 3), dtype=torch.float32, device=dev)
    acc_depth = torch.empty((Hh, Wh, 3), dtype=torch.float32, device=dev)
    st = stream_ptr()
    for i in range(P):
        d_i = depths[i:i + 1]
        check(lib.msi_highres_plane(ptr(ref), ptr(src), dt, 1, ptr(poses), ptr(baseline), ptr(d_i), *tb.ptrs(), Hh, Wh,
                                    ptr(bw), ptr(al), lh, lw, L, i, ptr(layer), st), "msi_highres_plane")
        check(lib.msi_highres_composite(ptr(layer), ptr(pose_rt), ptr(pos), ptr(d_i), *tb.ptrs(), Hh, Wh, i, P,
                                        ptr(acc_rgb), ptr(acc_depth), st), "msi_highres_composite")
    return acc_rgb, acc_depth


def deprocess_high_res(hres_output, hres_depth):
    """test.py:384-386 + utils.write_image: ((x+1)/2)*255 resp. d*255, clipped to [0,255], uint8."""
    out = (((hres_output + 1.0) / 2.0) * 255.0).clamp(0, 255).to(torch.uint8)
    dep = (hres_depth * 255.0).clamp(0, 255).to(torch.uint8)
    return out, dep
