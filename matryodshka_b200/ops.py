"""Tensor-level entry points over the C ABI: one Python function per kernel family.

Inputs and outputs are torch CUDA tensors (torch is the allocator and the stream owner,
nothing else); the arithmetic happens in libmsi_b200.so.  Work is enqueued on torch's current
stream.  The ERP-coord tables (cos/sin of the pixel-centre longitudes/latitudes) are computed
once per (H, W, device) on the host in float32 -- the same way a CPU evaluation of the
reference would -- and uploaded, which is what makes the project_ods `disc < 0` mask
bit-reproducible on the GPU (csrc/geom_device.cuh).
"""
from __future__ import annotations

import ctypes
import functools

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _linspace_tf(start, stop, num):
    """[TF-1.14 LinSpace]: float32 ``start + step * i`` with ``step = (stop - start) / (num - 1)``."""
    start = np.float32(start)
    stop = np.float32(stop)
    if num == 1:
        return np.array([start], np.float32)
    step = np.float32((stop - start) / np.float32(num - 1))
    return (start + step * np.arange(num).astype(np.float32)).astype(np.float32)


def lat_long_axes(H, W):
    """Pixel-centre longitudes s[W] and latitudes t[H] (spherical.py:42-44), float32 NumPy."""
    s = _linspace_tf(-np.pi + np.pi / W, np.pi - np.pi / W, W)
    t = _linspace_tf(-np.pi / 2.0 + np.pi / (2 * H), np.pi / 2.0 - np.pi / (2 * H), H)
    return s, t


class ErpTables:
    """cos/sin of the ERP pixel-centre angles on a device."""

    def __init__(self, H, W, device):
        s, t = lat_long_axes(H, W)
        self.H, self.W = H, W
        self.s, self.t = s, t
        self.cos_s = torch.from_numpy(np.cos(s)).to(device)
        self.sin_s = torch.from_numpy(np.sin(s)).to(device)
        self.cos_t = torch.from_numpy(np.cos(t)).to(device)
        self.sin_t = torch.from_numpy(np.sin(t)).to(device)

    def ptrs(self):
        return ptr(self.cos_s), ptr(self.sin_s), ptr(self.cos_t), ptr(self.sin_t)


@functools.lru_cache(maxsize=32)
def _tables_cached(H, W, dev_str):
    return ErpTables(H, W, torch.device(dev_str))


def erp_tables(H, W, device) -> ErpTables:
    return _tables_cached(int(H), int(W), str(torch.device(device)))


def _dev_f32(x, device, shape=None):
    """Small host array / tensor -> contiguous float32 CUDA tensor."""
    _lib.require_cuda()   # (the mirror API reaches the kernels through torch.ops.msi.*: fail with our own message first)
    if torch.is_tensor(x):
        t = x.to(device=device, dtype=torch.float32)
    else:
        t = torch.as_tensor(np.asarray(x, dtype=np.float32), device=device)
    if shape is not None:
        t = t.reshape(shape)
    return t.contiguous()


class SweepTable:
    """Cached sweep coordinates of one rig (msi_sweep_table_build): ``table`` [frames,H,W,P,4] float32 =
    (u_ref, v_ref, u_src, v_src) per (pixel, plane); frames = 1 when every frame of the batch has the same
    eye poses and baseline, else B."""

    def __init__(self, table, frames, H, W, P):
        self.table, self.frames, self.H, self.W, self.P = table, frames, H, W, P


_SWEEP_TABLES = {}        # key -> SweepTable, insertion-ordered (oldest first)
_SWEEP_TABLES_MAX = 4     # a table is H*W*P*16 bytes (105 MB at 320x640x32)


def sweep_table(poses, baselines, depths, H, W, device) -> SweepTable:
    """The coordinate table for (poses [B,2,4,4], baselines [B], depths [P]) on ``device``, built on torch's
    current stream the first time a rig is seen and then served from a small LRU cache keyed by the VALUES
    (float32 bits) of the rig -- any change of a pose, a baseline, the depths or the image size is a new
    table.  Host arrays in, so the key costs no device synchronisation."""
    _lib.require_cuda()
    lib = _lib.load()
    device = torch.device(device)
    poses = np.ascontiguousarray(np.asarray(poses, np.float32).reshape(-1, 2, 16))
    baselines = np.ascontiguousarray(np.asarray(baselines, np.float32).reshape(-1))
    depths = np.ascontiguousarray(np.asarray(depths, np.float32).reshape(-1))
    B, P = poses.shape[0], depths.size
    assert baselines.size == B
    if B > 1 and (poses == poses[:1]).all() and (baselines == baselines[0]).all():
        poses, baselines = poses[:1], baselines[:1]          # one rig for the whole batch
    frames = poses.shape[0]
    key = (poses.tobytes(), baselines.tobytes(), depths.tobytes(), int(H), int(W), str(device))
    hit = _SWEEP_TABLES.pop(key, None)
    if hit is None:
        table = torch.empty((frames, H, W, P, 4), dtype=torch.float32, device=device)
        tb = erp_tables(H, W, device)
        d_poses, d_base, d_depths = (torch.from_numpy(a).to(device) for a in (poses, baselines, depths))
        check(lib.msi_sweep_table_build(ptr(d_poses), ptr(d_base), ptr(d_depths), *tb.ptrs(), frames, H, W, P,
                                        ptr(table), stream_ptr()), "msi_sweep_table_build")
        # later users may run on other streams (frame lanes): the table is complete before it is handed out
        torch.cuda.current_stream(device).synchronize()
        hit = SweepTable(table, frames, H, W, P)
        while len(_SWEEP_TABLES) >= _SWEEP_TABLES_MAX:
            _SWEEP_TABLES.pop(next(iter(_SWEEP_TABLES)))
    _SWEEP_TABLES[key] = hit
    return hit


def psv_gather(ref, src, table: SweepTable, *, preprocess=True, want_f32=True, hi_lo=None, c_stride=None,
               use_scratch=True, scratch=None):
    """msi_psv_gather: the plane-sweep volume from cached coordinates; same outputs (same bits) as psv_build."""
    _lib.require_cuda()
    lib = _lib.load()
    assert ref.shape == src.shape and ref.dim() == 4 and ref.shape[3] == 3
    B, H, W, _ = ref.shape
    assert (table.H, table.W) == (H, W) and table.frames in (1, B), "sweep table built for another shape"
    P = table.P
    dev = ref.device
    if ref.dtype == torch.uint8:
        dt = _lib.IMG_U8
    elif ref.dtype == torch.float32:
        dt = _lib.IMG_F32
    else:
        raise _lib.MsiError(f"psv_gather: unsupported image dtype {ref.dtype}")
    out = torch.empty((B, H, W, 6 * P), dtype=torch.float32, device=dev) if want_f32 else None
    hi = lo = None
    cs = 6 * P
    if hi_lo is not None:
        hi, lo = hi_lo
        cs = int(c_stride if c_stride is not None else hi.shape[-1])
    if scratch is None and use_scratch:
        scratch = psv_scratch(B, H, W, dev)
    check(lib.msi_psv_gather(ptr(ref.contiguous()), ptr(src.contiguous()), dt, 1 if preprocess else 0, ptr(table.table),
                             table.frames, B, H, W, P, ptr(out), ptr(hi), ptr(lo), cs, ptr(scratch),
                             scratch.numel() if scratch is not None else 0, stream_ptr()), "msi_psv_gather")
    return out


def psv_build(ref, src, poses, baselines, depths, *, preprocess=True, want_f32=True, hi_lo=None, c_stride=None,
              use_scratch=True, cache_coords=False):
    """msi_psv_build.  ref/src: [B,H,W,3] float32 or uint8 CUDA tensors; poses [B,2,4,4];
    baselines [B]; depths [P].  Returns the float32 PSV [B,H,W,6P] (or None); ``hi_lo`` is an
    optional (hi, lo) pair of fp16 [B,H,W,c_stride] tensors filled with the conv-operand copy.
    ``cache_coords``: take the sample coordinates from the per-rig table (sweep_table / psv_gather; host
    ``poses`` / ``baselines`` only) -- same bits, a pure gather from the second frame of a rig on."""
    _lib.require_cuda()
    lib = _lib.load()
    assert ref.shape == src.shape and ref.dim() == 4 and ref.shape[3] == 3
    B, H, W, _ = ref.shape
    dev = ref.device
    if cache_coords and not torch.is_tensor(poses) and not torch.is_tensor(baselines) and not torch.is_tensor(depths):
        tbl = sweep_table(np.asarray(poses, np.float32).reshape(B, 2, 16), baselines, depths, H, W, dev)
        return psv_gather(ref, src, tbl, preprocess=preprocess, want_f32=want_f32, hi_lo=hi_lo, c_stride=c_stride,
                          use_scratch=use_scratch)
    depths = _dev_f32(depths, dev, (-1,))
    P = depths.numel()
    poses = _dev_f32(poses, dev, (B, 2, 16))
    baselines = _dev_f32(baselines, dev, (B,))
    tb = erp_tables(H, W, dev)
    if ref.dtype == torch.uint8:
        dt = _lib.IMG_U8
    elif ref.dtype == torch.float32:
        dt = _lib.IMG_F32
    else:
        raise _lib.MsiError(f"psv_build: unsupported image dtype {ref.dtype}")
    out = torch.empty((B, H, W, 6 * P), dtype=torch.float32, device=dev) if want_f32 else None
    hi = lo = None
    cs = 6 * P
    if hi_lo is not None:
        hi, lo = hi_lo
        cs = int(c_stride if c_stride is not None else hi.shape[-1])
    scratch = psv_scratch(B, H, W, dev) if use_scratch else None
    check(lib.msi_psv_build(ptr(ref.contiguous()), ptr(src.contiguous()), dt, 1 if preprocess else 0,
                            ptr(poses), ptr(baselines), ptr(depths), *tb.ptrs(), B, H, W, P,
                            ptr(out), ptr(hi), ptr(lo), cs, ptr(scratch), scratch.numel() if scratch is not None else 0,
                            stream_ptr()), "msi_psv_build")
    return out


def psv_scratch(B, H, W, device):
    """16-byte aligned scratch for the two-eye form of msi_psv_build."""
    n = int(_lib.load().msi_psv_scratch_bytes(B, H, W))
    return torch.empty(n, dtype=torch.uint8, device=device)


def sweep_coords(poses, baselines, depths, B, H, W, device):
    """msi_sweep_coords -> (uv [B,2,P,H,W,2] float32, valid [B,2,P,H,W] uint8)."""
    _lib.require_cuda()
    lib = _lib.load()
    depths = _dev_f32(depths, device, (-1,))
    P = depths.numel()
    poses = _dev_f32(poses, device, (B, 2, 16))
    baselines = _dev_f32(baselines, device, (B,))
    tb = erp_tables(H, W, device)
    uv = torch.empty((B, 2, P, H, W, 2), dtype=torch.float32, device=device)
    valid = torch.empty((B, 2, P, H, W), dtype=torch.uint8, device=device)
    check(lib.msi_sweep_coords(ptr(poses), ptr(baselines), ptr(depths), *tb.ptrs(), B, H, W, P, ptr(uv), ptr(valid),
                               stream_ptr()), "msi_sweep_coords")
    return uv, valid


def color_pred_channels(which_color_pred, num_msi_planes):
    """Output channels of the net for a colour-prediction scheme (msi.py:107-116)."""
    L = num_msi_planes
    return {"blend_psv": 2 * L, "blend_bg": 2 * L + 3, "blend_bg_psv": 3 * L + 3, "alpha_only": L}[which_color_pred]


def _pixel_stride(pred):
    """Floats between consecutive pixels of a [B,H,W,C] prediction that may be a channel slice of a wider,
    contiguous buffer (the tensor-core net pads its head to a multiple of 64 channels); None if it is not."""
    B, H, W, C = pred.shape
    st = pred.stride()
    if st[3] == 1 and st[2] >= C and st[1] == W * st[2] and st[0] == H * st[1]:
        return st[2]
    return None


def rgba_assemble_ex(pred, psv=None, which_color_pred="blend_psv", num_msi_planes=None, *, want_weights=False,
                     hi_lo=None, c_stride=None, out=None):
    """msi_rgba_assemble_strided: RGBA layers for any `which_color_pred` (msi.py:117-268).  pred [B,H,W,n_pred]
    (read in place when it is a channel slice of the net's padded output); psv float32 [B,H,W,6L] or the fp16
    (hi, lo) operand pair.  Returns (rgba, blend_weights, alphas, bg_blend_weights)."""
    _lib.require_cuda()
    lib = _lib.load()
    B, H, W, n_pred = pred.shape
    L = num_msi_planes
    assert n_pred == color_pred_channels(which_color_pred, L), (n_pred, which_color_pred, L)
    dev = pred.device
    stride = _pixel_stride(pred)
    if stride is None:
        pred, stride = pred.contiguous(), n_pred
    hi = lo = None
    cs = 6 * L
    if psv is None:
        hi, lo = hi_lo
        cs = int(c_stride if c_stride is not None else hi.shape[-1])
    else:
        assert psv.shape[-1] == 6 * L, "the colour schemes need num_psv_planes == num_msi_planes"
        psv = psv.contiguous()
    mk = lambda: torch.empty((B, H, W, L), dtype=torch.float32, device=dev)
    rgba = out if out is not None else torch.empty((B, H, W, L, 4), dtype=torch.float32, device=dev)
    bw = mk() if want_weights and which_color_pred != "alpha_only" else None
    al = mk() if want_weights else None
    bgw = mk() if want_weights and which_color_pred == "blend_bg_psv" else None
    check(lib.msi_rgba_assemble_strided(ctypes.c_void_p(pred.data_ptr()), n_pred, stride, ptr(psv), ptr(hi), ptr(lo), cs, B, H, W, L,
                                        _lib.COLOR_MODES[which_color_pred], ptr(rgba), ptr(bw), ptr(al), ptr(bgw),
                                        stream_ptr()), "msi_rgba_assemble_strided")
    return rgba, bw, al, bgw


def rgba_assemble(pred, psv=None, *, hi_lo=None, c_stride=None, want_weights=False):
    """msi_rgba_assemble.  pred [B,H,W,2L]; psv float32 [B,H,W,6L] or the fp16 (hi, lo) pair.
    Returns (rgba [B,H,W,L,4], blend_weights | None, alphas | None)."""
    _lib.require_cuda()
    lib = _lib.load()
    B, H, W, C2 = pred.shape
    L = C2 // 2
    dev = pred.device
    rgba = torch.empty((B, H, W, L, 4), dtype=torch.float32, device=dev)
    bw = torch.empty((B, H, W, L), dtype=torch.float32, device=dev) if want_weights else None
    al = torch.empty((B, H, W, L), dtype=torch.float32, device=dev) if want_weights else None
    hi = lo = None
    cs = 6 * L
    if psv is None:
        hi, lo = hi_lo
        cs = int(c_stride if c_stride is not None else hi.shape[-1])
    else:
        assert psv.shape[-1] == 6 * L, "blend_psv needs num_psv_planes == num_msi_planes"
    check(lib.msi_rgba_assemble(ptr(pred.contiguous()), ptr(psv.contiguous() if psv is not None else None),
                                ptr(hi), ptr(lo), cs, B, H, W, L, ptr(rgba), ptr(bw), ptr(al), stream_ptr()),
          "msi_rgba_assemble")
    return rgba, bw, al


def render_composite(rgba, tgt_pose_rt, tgt_pos, depths, *, want_rgb=True, want_depth=True, want_u8=True,
                     out=None):
    """msi_render_composite.  rgba [B,H,W,L,4]; tgt_pose_rt [B,4,4]; tgt_pos [B,3]; depths [L].
    Returns dict with 'rgb', 'depth' (float32 [B,H,W,3]) and 'rgb_u8', 'depth_u8'."""
    _lib.require_cuda()
    lib = _lib.load()
    B, H, W, L, four = rgba.shape
    assert four == 4
    dev = rgba.device
    depths = _dev_f32(depths, dev, (-1,))
    assert depths.numel() == L
    pose = _dev_f32(tgt_pose_rt, dev, (B, 16))
    pos = _dev_f32(tgt_pos, dev, (B, 3))
    tb = erp_tables(H, W, dev)
    res = out if out is not None else {}
    if want_rgb and "rgb" not in res:
        res["rgb"] = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
    if want_depth and "depth" not in res:
        res["depth"] = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
    if want_u8 and want_rgb and "rgb_u8" not in res:
        res["rgb_u8"] = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    if want_u8 and want_depth and "depth_u8" not in res:
        res["depth_u8"] = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)
    check(lib.msi_render_composite(ptr(rgba.contiguous()), ptr(pose), ptr(pos), ptr(depths), *tb.ptrs(), B, H, W, L,
                                   ptr(res.get("rgb")), ptr(res.get("depth")), ptr(res.get("rgb_u8")),
                                   ptr(res.get("depth_u8")), stream_ptr()), "msi_render_composite")
    return res


def render_ods(rgba, pose_rt, order, baselines, depths, *, want_u8=False):
    """msi_render_ods: the MSI [B,H,W,L,4] seen from one ODS eye (order +1 / -1) under pose_rt
    [B,4,4].  Returns rgb [B,H,W,3] float32 (and the uint8 deprocessed image when want_u8)."""
    _lib.require_cuda()
    lib = _lib.load()
    B, H, W, L, four = rgba.shape
    assert four == 4
    dev = rgba.device
    depths = _dev_f32(depths, dev, (-1,))
    assert depths.numel() == L
    pose = _dev_f32(pose_rt, dev, (B, 16))
    base = _dev_f32(baselines, dev, (B,))
    tb = erp_tables(H, W, dev)
    rgb = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
    u8 = torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev) if want_u8 else None
    check(lib.msi_render_ods(ptr(rgba.contiguous()), ptr(pose), float(order), ptr(base), ptr(depths), *tb.ptrs(),
                             B, H, W, L, ptr(rgb), ptr(u8), stream_ptr()), "msi_render_ods")
    return (rgb, u8) if want_u8 else rgb


def viewing_window_pose(viewing_window, B):
    """projector.py:80-85: rotation by viewing_window * pi / 2 about y, as a [B,4,4] float32 [R|0] pose.
    [tensorflow_graphics 1.0.0 rotation_matrix_3d.from_euler, R = Rz Ry Rx, restated for angles (0, a, 0):
    float32 sin / cos of the float32 angle.]"""
    a = np.float32(viewing_window * np.pi / 2.0)
    sy, cy = np.sin(a, dtype=np.float32), np.cos(a, dtype=np.float32)
    m = np.array([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1]], dtype=np.float32)
    return np.tile(m[None], (B, 1, 1))


def render_perspective(rgba, tgt_pos, depths, *, viewing_window=3, psp_height=270, psp_width=480, want_u8=False):
    """msi_render_perspective: pinhole view [B,psp_height,psp_width,3] of an MSI (msi.py:475-500)."""
    _lib.require_cuda()
    lib = _lib.load()
    B, H, W, L, _ = rgba.shape
    dev = rgba.device
    pose = _dev_f32(viewing_window_pose(viewing_window, B), dev).reshape(B, 16).contiguous()
    pos = _dev_f32(tgt_pos, dev).reshape(B, 3).contiguous()
    d = _dev_f32(np.asarray(depths, dtype=np.float32), dev)
    s_axis = _dev_f32(_linspace_tf(-1.0 + 1.0 / psp_width, 1.0 - 1.0 / psp_width, psp_width), dev)   # spherical.py:46-48
    t_axis = _dev_f32(_linspace_tf(-1.0 + 1.0 / psp_height, 1.0 - 1.0 / psp_height, psp_height), dev)
    out = torch.empty((B, psp_height, psp_width, 3), dtype=torch.float32, device=dev)
    u8 = torch.empty((B, psp_height, psp_width, 3), dtype=torch.uint8, device=dev) if want_u8 else None
    check(lib.msi_render_perspective(ptr(rgba.contiguous()), ptr(pose), ptr(pos), ptr(d), ptr(s_axis), ptr(t_axis),
                                     B, H, W, L, psp_height, psp_width, ptr(out), ptr(u8), stream_ptr()),
          "msi_render_perspective")
    return (out, u8) if want_u8 else out


def intersect_sphere_coords(tgt_pose_rt, tgt_pos, depths, B, H, W, device, fast=False):
    """msi_intersect_sphere_coords[_ex] -> uv [B,L,H,W,2]; ``fast``: the coordinates the fused render kernel samples at."""
    _lib.require_cuda()
    lib = _lib.load()
    depths = _dev_f32(depths, device, (-1,))
    L = depths.numel()
    pose = _dev_f32(tgt_pose_rt, device, (B, 16))
    pos = _dev_f32(tgt_pos, device, (B, 3))
    tb = erp_tables(H, W, device)
    uv = torch.empty((B, L, H, W, 2), dtype=torch.float32, device=device)
    check(lib.msi_intersect_sphere_coords_ex(ptr(pose), ptr(pos), ptr(depths), *tb.ptrs(), B, H, W, L, 1 if fast else 0,
                                             ptr(uv), stream_ptr()), "msi_intersect_sphere_coords_ex")
    return uv


def project_layers(rgba, tgt_pose_rt, tgt_pos, depths):
    """msi_project_layers -> [L,B,H,W,4]."""
    _lib.require_cuda()
    lib = _lib.load()
    B, H, W, L, _ = rgba.shape
    dev = rgba.device
    depths = _dev_f32(depths, dev, (-1,))
    pose = _dev_f32(tgt_pose_rt, dev, (B, 16))
    pos = _dev_f32(tgt_pos, dev, (B, 3))
    tb = erp_tables(H, W, dev)
    out = torch.empty((L, B, H, W, 4), dtype=torch.float32, device=dev)
    check(lib.msi_project_layers(ptr(rgba.contiguous()), ptr(pose), ptr(pos), ptr(depths), *tb.ptrs(), B, H, W, L,
                                 ptr(out), stream_ptr()), "msi_project_layers")
    return out


def resample(image, coords):
    """msi_resample (sampling.py:135-197).  image [N,H,W,C], coords [N,h,w,2] -> [N,h,w,C]."""
    _lib.require_cuda()
    lib = _lib.load()
    N, H, W, C = image.shape
    n2, h, w, two = coords.shape
    assert n2 == N and two == 2
    out = torch.empty((N, h, w, C), dtype=torch.float32, device=image.device)
    check(lib.msi_resample(ptr(image.contiguous().float()), ptr(coords.contiguous().float()), N, H, W, C, h, w,
                           ptr(out), stream_ptr()), "msi_resample")
    return out


def over_composite(layers, depth_mode=False):
    """msi_over_composite.  layers [L,B,H,W,4] back to front -> [B,H,W,3]."""
    _lib.require_cuda()
    lib = _lib.load()
    L, B, H, W, four = layers.shape
    assert four == 4
    out = torch.empty((B, H, W, 3), dtype=torch.float32, device=layers.device)
    check(lib.msi_over_composite(ptr(layers.contiguous()), L, B, H, W, 1 if depth_mode else 0, ptr(out),
                                 stream_ptr()), "msi_over_composite")
    return out


OP_BACKPROJECT_SPHERICAL, OP_APPLY_POSE, OP_PROJECT_ODS, OP_PROJECT_SPHERICAL, OP_THETA_PHI_TO_PIXELS = range(5)


def point_op(op, a, b, c=None, *, planes=1, pose=None, order=1.0, baseline=0.0, H=2, W=2, want_valid=False):
    """msi_point_op: the point-wise geometry/spherical.py functions on flat float32 CUDA tensors.
    Returns (o0, o1, o2) for back-projection / apply_pose, uv [n,2] (and valid) for the projections."""
    _lib.require_cuda()
    lib = _lib.load()
    a = a.contiguous().float()
    b = b.contiguous().float()
    dev = a.device
    c = c.contiguous().float() if c is not None else None
    if op == OP_BACKPROJECT_SPHERICAL:
        n = a.numel()
        planes = c.numel()
        outs = [torch.empty((planes, n), dtype=torch.float32, device=dev) for _ in range(3)]
    elif op == OP_APPLY_POSE:
        planes = a.shape[0]
        n = a.numel() // planes
        outs = [torch.empty_like(a) for _ in range(3)]
    else:
        n = a.numel()
        outs = [torch.empty((n, 2), dtype=torch.float32, device=dev), None, None]
    pose_t, per_plane = None, 0
    if pose is not None:
        pose_t = _dev_f32(pose, dev).reshape(-1, 16).contiguous()
        per_plane = 1 if pose_t.shape[0] > 1 else 0
    valid = torch.empty((n,), dtype=torch.uint8, device=dev) if want_valid else None
    check(lib.msi_point_op(op, ptr(a), ptr(b), ptr(c), n, planes, ptr(pose_t), per_plane, float(order), float(baseline),
                           int(H), int(W), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(valid), stream_ptr()),
          "msi_point_op")
    if op <= OP_APPLY_POSE:
        return tuple(outs)
    return (outs[0], valid) if want_valid else outs[0]
