"""PyTorch custom-op registration of the C ABI (north_star: "Python host code binding hand-written
sm_100a CUDA kernels through PyTorch custom ops").

Every op below is a ``torch.library.custom_op`` in the namespace ``msi`` whose CUDA implementation calls one entry
point of ``include/msi_b200.h`` (through ``ops.py`` -> ctypes -> ``libmsi_b200.so``) on torch's current stream; the
outputs are allocator-owned tensors returned by the dispatcher, and every op carries a fake (meta) implementation for
shape inference, so the ops trace under FakeTensorMode / ``torch.compile`` as opaque nodes.  There is NO CPU kernel:
calling an op on CPU tensors raises ``NotImplementedError`` from the dispatcher (no fallback).

    torch.ops.msi.psv_build(ref, src, poses, baselines, depths, preprocess)         -> psv   [B,H,W,6P]
    torch.ops.msi.sweep_table(poses, baselines, depths, H, W)                       -> table [F,H,W,P,4]
    torch.ops.msi.psv_gather(ref, src, table, preprocess)                           -> psv   [B,H,W,6P]
    torch.ops.msi.net_forward(psv, engine)                                          -> pred  [B,H,W,c_out]
    torch.ops.msi.net_forward_rgba(psv, engine)                                     -> rgba  [B,H,W,L,4]
    torch.ops.msi.rgba_assemble(pred, psv, mode, L)                                 -> rgba, blend_weights, alphas, bg_blend_weights
    torch.ops.msi.render_composite(rgba, tgt_pose_rt, tgt_pos, depths)              -> rgb, depth, rgb_u8, depth_u8
    torch.ops.msi.project_layers(rgba, tgt_pose_rt, tgt_pos, depths)                -> layers [L,B,H,W,4]
    torch.ops.msi.intersect_sphere_coords(tgt_pose_rt, tgt_pos, depths, H, W, fast) -> uv [B,L,H,W,2]
    torch.ops.msi.resample(image, coords)                                           -> [N,h,w,C]
    torch.ops.msi.over_composite(layers, depth_mode)                                -> [B,H,W,3]

``engine`` is an integer handle of a ``runtime.NetEngine`` (a net with its packed weights and workspace is a
stateful object; ``register_engine`` / ``release_engine`` manage the handles).  The mirror API (``msi.MSI``,
``geometry.*``) calls these ops; the graph-captured ``runtime.MSIPipeline`` calls the C ABI directly inside its
capture (a CUDA graph replays kernels, not dispatcher calls).

Replaces (reference file:line): the graphs of stock TF ops behind ``MSI.format_network_input`` (msi.py:1094-1130),
``nets.msi_coord_train_net`` (nets.py:471-515), the ``infer_msi`` layer assembly (msi.py:117-273),
``MSI.msi_render_equirect_view`` / ``_depth`` (msi.py:384-429), ``projector.projective_forward_sphere``
(projector.py:34-62), ``spherical.intersect_sphere`` (spherical.py:268-326), ``sampling.resample``
(sampling.py:135-197) and ``projector.over_composite[_depth]`` (projector.py:225-265).
"""
from __future__ import annotations

import itertools
import weakref
from typing import Tuple

import torch
from torch import Tensor
from torch.library import custom_op

from . import ops

_MODES = ("blend_psv", "blend_bg", "blend_bg_psv", "alpha_only")   # = MSI_COLOR_* of include/msi_b200.h

# ---- engine handles ------------------------------------------------------------------------------------------
_engines = {}
_next_handle = itertools.count(1)


def register_engine(engine) -> int:
    """Handle for ``torch.ops.msi.net_forward*``; the registry keeps a weak reference (the caller owns the engine)."""
    h = next(_next_handle)
    _engines[h] = weakref.ref(engine)
    return h


def release_engine(handle: int) -> None:
    _engines.pop(int(handle), None)


def _engine(handle: int):
    ref = _engines.get(int(handle))
    eng = ref() if ref is not None else None
    if eng is None:
        raise RuntimeError(f"msi: no live NetEngine behind handle {handle}")
    return eng


def engine_shape(handle: int) -> Tuple[int, int, int, int]:
    e = _engine(handle)
    return e.H, e.W, e.c_in, e.c_out


# ---- stage 1 ---------------------------------------------------------------------------------------------------
@custom_op("msi::psv_build", mutates_args=(), device_types="cuda")
def psv_build(ref: Tensor, src: Tensor, poses: Tensor, baselines: Tensor, depths: Tensor, preprocess: bool) -> Tensor:
    return ops.psv_build(ref, src, poses, baselines, depths, preprocess=preprocess)


@psv_build.register_fake
def _(ref, src, poses, baselines, depths, preprocess):
    B, H, W, _ = ref.shape
    return ref.new_empty((B, H, W, 6 * depths.numel()), dtype=torch.float32)


@custom_op("msi::sweep_table", mutates_args=(), device_types="cuda")
def sweep_table(poses: Tensor, baselines: Tensor, depths: Tensor, H: int, W: int) -> Tensor:
    # (the op form builds a fresh table; ops.sweep_table adds the per-rig cache on host arrays)
    from ._lib import check, load, ptr, stream_ptr
    frames, P = poses.shape[0], depths.numel()
    table = torch.empty((frames, H, W, P, 4), dtype=torch.float32, device=depths.device)
    tb = ops.erp_tables(H, W, depths.device)
    check(load().msi_sweep_table_build(ptr(poses.reshape(frames, 2, 16).float().contiguous()), ptr(baselines.float().contiguous()),
                                       ptr(depths.float().contiguous()), *tb.ptrs(), frames, H, W, P, ptr(table),
                                       stream_ptr()), "msi_sweep_table_build")
    return table


@sweep_table.register_fake
def _(poses, baselines, depths, H, W):
    return depths.new_empty((poses.shape[0], H, W, depths.numel(), 4), dtype=torch.float32)


@custom_op("msi::psv_gather", mutates_args=(), device_types="cuda")
def psv_gather(ref: Tensor, src: Tensor, table: Tensor, preprocess: bool) -> Tensor:
    F, H, W, P, _ = table.shape
    return ops.psv_gather(ref, src, ops.SweepTable(table.contiguous(), F, H, W, P), preprocess=preprocess)


@psv_gather.register_fake
def _(ref, src, table, preprocess):
    B, H, W, _ = ref.shape
    return ref.new_empty((B, H, W, 6 * table.shape[3]), dtype=torch.float32)


# ---- stage 2 ---------------------------------------------------------------------------------------------------
@custom_op("msi::net_forward", mutates_args=(), device_types="cuda")
def net_forward(psv: Tensor, engine: int) -> Tensor:
    return _engine(engine).forward(psv).contiguous()


@net_forward.register_fake
def _(psv, engine):
    B, H, W, _ = psv.shape
    return psv.new_empty((B, H, W, engine_shape(engine)[3]), dtype=torch.float32)


@custom_op("msi::net_forward_rgba", mutates_args=(), device_types="cuda")
def net_forward_rgba(psv: Tensor, engine: int) -> Tensor:
    return _engine(engine).forward_rgba(psv)


@net_forward_rgba.register_fake
def _(psv, engine):
    B, H, W, _ = psv.shape
    return psv.new_empty((B, H, W, engine_shape(engine)[3] // 2, 4), dtype=torch.float32)


@custom_op("msi::rgba_assemble", mutates_args=(), device_types="cuda")
def rgba_assemble(pred: Tensor, psv: Tensor, mode: int, L: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    rgba, bw, al, bgw = ops.rgba_assemble_ex(pred, psv, _MODES[mode], L, want_weights=True)
    empty = pred.new_empty((0,))
    return rgba, (bw if bw is not None else empty), al, (bgw if bgw is not None else empty)


@rgba_assemble.register_fake
def _(pred, psv, mode, L):
    B, H, W, _ = pred.shape
    full, empty = pred.new_empty((B, H, W, L)), pred.new_empty((0,))
    return (pred.new_empty((B, H, W, L, 4)), empty if _MODES[mode] == "alpha_only" else full, pred.new_empty((B, H, W, L)),
            pred.new_empty((B, H, W, L)) if _MODES[mode] == "blend_bg_psv" else empty)


# ---- stage 3 ---------------------------------------------------------------------------------------------------
@custom_op("msi::render_composite", mutates_args=(), device_types="cuda")
def render_composite(rgba: Tensor, tgt_pose_rt: Tensor, tgt_pos: Tensor, depths: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    res = ops.render_composite(rgba, tgt_pose_rt, tgt_pos, depths)
    return res["rgb"], res["depth"], res["rgb_u8"], res["depth_u8"]


@render_composite.register_fake
def _(rgba, tgt_pose_rt, tgt_pos, depths):
    B, H, W = rgba.shape[:3]
    f = lambda dt: rgba.new_empty((B, H, W, 3), dtype=dt)  # noqa: E731
    return f(torch.float32), f(torch.float32), f(torch.uint8), f(torch.uint8)


@custom_op("msi::project_layers", mutates_args=(), device_types="cuda")
def project_layers(rgba: Tensor, tgt_pose_rt: Tensor, tgt_pos: Tensor, depths: Tensor) -> Tensor:
    return ops.project_layers(rgba, tgt_pose_rt, tgt_pos, depths)


@project_layers.register_fake
def _(rgba, tgt_pose_rt, tgt_pos, depths):
    B, H, W, L, _ = rgba.shape
    return rgba.new_empty((L, B, H, W, 4))


@custom_op("msi::intersect_sphere_coords", mutates_args=(), device_types="cuda")
def intersect_sphere_coords(tgt_pose_rt: Tensor, tgt_pos: Tensor, depths: Tensor, H: int, W: int, fast: bool) -> Tensor:
    B = tgt_pos.reshape(-1, 3).shape[0]
    return ops.intersect_sphere_coords(tgt_pose_rt, tgt_pos, depths, B, H, W, depths.device, fast=fast)


@intersect_sphere_coords.register_fake
def _(tgt_pose_rt, tgt_pos, depths, H, W, fast):
    return depths.new_empty((tgt_pos.reshape(-1, 3).shape[0], depths.numel(), H, W, 2), dtype=torch.float32)


@custom_op("msi::resample", mutates_args=(), device_types="cuda")
def resample(image: Tensor, coords: Tensor) -> Tensor:
    return ops.resample(image, coords)


@resample.register_fake
def _(image, coords):
    N, h, w, _ = coords.shape
    return image.new_empty((N, h, w, image.shape[3]), dtype=torch.float32)


@custom_op("msi::over_composite", mutates_args=(), device_types="cuda")
def over_composite(layers: Tensor, depth_mode: bool) -> Tensor:
    return ops.over_composite(layers, depth_mode=depth_mode)


@over_composite.register_fake
def _(layers, depth_mode):
    L, B, H, W, _ = layers.shape
    return layers.new_empty((B, H, W, 3))


OP_NAMES = ("psv_build", "sweep_table", "psv_gather", "net_forward", "net_forward_rgba", "rgba_assemble",
            "render_composite", "project_layers", "intersect_sphere_coords", "resample", "over_composite")
