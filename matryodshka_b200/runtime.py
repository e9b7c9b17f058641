"""Runtime objects of the MSI inference path: the conv-net engine and the frame pipeline.

``NetEngine``   owns one ``msi_net`` (C ABI), its workspace / parameter arena (torch uint8
                CUDA storage) and the packed weights.
``MSIPipeline`` runs whole frames: host images -> PSV -> net -> RGBA layers -> rendered ERP
                view (+ depth), with the per-batch launch sequence captured in a CUDA graph,
                pinned host staging for the end-to-end path, and frame sharding across ranks
                with one all-gather of the rendered outputs.
"""
from __future__ import annotations

import ctypes
from ctypes import c_void_p
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr
from .nets import ARCH

_CONV_IMPL = {"tcgen05": _lib.CONV_TCGEN05, "simt": _lib.CONV_SIMT}
# "fp16_fp8x": fp16 main product + e4m3 cross terms on the N = 128 layers (2 MMA units per product instead of 3)
_PRECISION = {"fp16x3": _lib.PREC_FP16X3, "fp16": _lib.PREC_FP16, "fp16_fp8x": _lib.PREC_FP16_FP8X}
_VARIANT = {"coord": _lib.NET_COORD, "wrap": _lib.NET_WRAP}


class nvtx_range:
    """NVTX range around one stage of the path (visible in Nsight Systems / ncu --nvtx; SURVEY.md 5 tracing row).
    Host-side markers only: they cost ~1 us and are harmless inside a CUDA-graph capture."""

    def __init__(self, name):
        self.name = "msi/" + name

    def __enter__(self):
        try:
            torch.cuda.nvtx.range_push(self.name)
            self._on = True
        except Exception:   # NVTX library unavailable: tracing is optional
            self._on = False
        return self

    def __exit__(self, *exc):
        if self._on:
            torch.cuda.nvtx.range_pop()
        return False


def _host_f32(x):
    return x.detach().cpu().numpy().astype(np.float32) if torch.is_tensor(x) else np.asarray(x, np.float32)


def _aligned_bytes(n, device):
    """uint8 CUDA buffer of >= n bytes whose base is 1024-byte aligned."""
    buf = torch.empty(n + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    return buf, off


class NetEngine:
    """nets.msi_coord_train_net (nets.py:471-515; ``variant="coord"``) or nets.msi_train_net
    (nets.py:387-469: no coord channel, circular-x / zero-y ``wrap_pad``; ``variant="wrap"``) on the GPU.

    weights: dict keyed by the TF checkpoint variable names (``net/<scope>/weights`` ...),
    NumPy arrays or tensors in the TF layouts."""

    _cache: Dict[tuple, "NetEngine"] = {}

    def __init__(self, weights, H, W, c_in, c_out, ngf=64, device="cuda", max_batch=1,
                 conv_impl="tcgen05", precision="fp16_fp8x", vscope="net", variant="coord"):
        _lib.require_cuda()
        self.lib = _lib.load()
        self.variant = variant
        self.device = torch.device(device)
        self.H, self.W, self.c_in, self.c_out, self.ngf = H, W, c_in, c_out, ngf
        # the tensor-core kernels tile Cout in 64s (the SIMT kernel in 4s): a head with 2L+3 / 3L+3 / L outputs
        # (blend_bg, blend_bg_psv, alpha_only) is padded with zero weights and the extra channels are dropped
        q = 64 if conv_impl == "tcgen05" else 4
        self.c_out_eng = -(-c_out // q) * q
        self.max_batch = max_batch
        self.conv_impl, self.precision = conv_impl, precision
        self._h = c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.msi_net_create_ex(ctypes.byref(self._h), H, W, c_in, self.c_out_eng, ngf, max_batch,
                                             _CONV_IMPL[conv_impl], _PRECISION[precision], _VARIANT[variant]),
                  "msi_net_create_ex")
            self.ws_bytes = int(self.lib.msi_net_workspace_bytes(self._h))
            self.arena_bytes = int(self.lib.msi_net_arena_bytes(self._h))
            self._ws, wo = _aligned_bytes(self.ws_bytes, self.device)
            self._arena, ao = _aligned_bytes(self.arena_bytes, self.device)
            check(self.lib.msi_net_bind(self._h, c_void_p(self._ws.data_ptr() + wo), self.ws_bytes,
                                        c_void_p(self._arena.data_ptr() + ao), self.arena_bytes), "msi_net_bind")
            self.in_c_stride = int(self.lib.msi_net_input_c_stride(self._h))
            self.load_weights(weights, vscope)
        self.launches_per_forward = int(self.lib.msi_net_num_launches_per_forward(self._h))

    _cache_max = 4   # every engine pins a workspace + arena (about 1 GB per frame of max_batch at 320 x 640)

    @staticmethod
    def _fingerprint(weights):
        """Cheap content key of a weights dict: names, shapes and a strided sample of every array.  (Keying on
        ``id(weights)`` would hand out stale packed weights after an in-place reload, or after CPython recycles the id.)"""
        items = []
        for name in sorted(weights):
            w = weights[name]
            a = w.detach().reshape(-1) if torch.is_tensor(w) else np.asarray(w).reshape(-1)
            step = max(1, a.shape[0] // 256)
            samp = a[::step][:256]
            samp = samp.float().cpu().numpy() if torch.is_tensor(samp) else np.asarray(samp, np.float32)
            items.append((name, tuple(w.shape), samp.tobytes()))
        return hash(tuple(items))

    @classmethod
    def cached(cls, weights, H, W, c_in, c_out, ngf, device, vscope="net", max_batch=1, **kw):
        """An engine for these weights and this shape out of a small LRU cache; an engine built for a larger batch is
        reused for any smaller one."""
        base = (cls._fingerprint(weights), H, W, c_in, c_out, ngf, str(device), vscope, tuple(sorted(kw.items())))
        for key in list(cls._cache):
            if key[:-1] == base and key[-1] >= max_batch:
                eng = cls._cache.pop(key)
                cls._cache[key] = eng          # most recently used last
                return eng
        eng = cls(weights, H, W, c_in, c_out, ngf, device, vscope=vscope, max_batch=max_batch, **kw)
        while len(cls._cache) >= cls._cache_max:
            cls._cache.pop(next(iter(cls._cache)))
        cls._cache[base + (max_batch,)] = eng
        return eng

    def load_weights(self, weights, vscope="net"):
        def dev(name):
            if name not in weights:
                return None
            w = weights[name]
            if not torch.is_tensor(w):
                w = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
            return w.to(device=self.device, dtype=torch.float32).contiguous()

        keep = []
        for l in ARCH:
            w = dev(f"{vscope}/{l.scope}/weights")
            if w is None:
                raise _lib.MsiError(f"missing weights for {vscope}/{l.scope}")
            g = dev(f"{vscope}/{l.scope}/LayerNorm/gamma")
            b = dev(f"{vscope}/{l.scope}/LayerNorm/beta")
            bias = dev(f"{vscope}/{l.scope}/biases")
            if l.scope == "color_pred" and self.c_out_eng != self.c_out:
                pad = self.c_out_eng - self.c_out
                w = torch.nn.functional.pad(w, (0, pad)).contiguous()        # [1,1,Cin,c_out] -> zero output channels
                bias = torch.nn.functional.pad(bias, (0, pad)).contiguous()
            keep += [w, g, b, bias]
            check(self.lib.msi_net_load_layer(self._h, l.scope.encode(), ptr(w), ptr(g), ptr(b), ptr(bias),
                                              stream_ptr()), f"msi_net_load_layer({l.scope})")
        torch.cuda.current_stream().synchronize()  # staging tensors in `keep` may now be freed

    def input_buffers(self, B):
        """(hi, lo) fp16 views [B,H,W,in_c_stride] of the workspace copy of the net input."""
        hi, lo = c_void_p(), c_void_p()
        check(self.lib.msi_net_input_buffers(self._h, ctypes.byref(hi), ctypes.byref(lo)), "msi_net_input_buffers")
        n = B * self.H * self.W * self.in_c_stride

        def view(p):
            off = p.value - self._ws.data_ptr()
            return self._ws[off:off + 2 * n].view(torch.float16).view(B, self.H, self.W, self.in_c_stride)
        return view(hi), view(lo)

    def forward(self, psv=None, *, hi_lo=None, out=None):
        """psv: float32 [B,H,W,c_in] (or hi_lo = the fp16 operand pair).  Returns pred [B,H,W,c_out]."""
        if psv is not None:
            B = psv.shape[0]
            assert tuple(psv.shape[1:]) == (self.H, self.W, self.c_in), psv.shape
            psv = psv.contiguous()
        else:
            B = hi_lo[0].shape[0]
        if out is None or out.shape[-1] != self.c_out_eng:
            out = torch.empty((B, self.H, self.W, self.c_out_eng), dtype=torch.float32, device=self.device)
        hi, lo = hi_lo if hi_lo is not None else (None, None)
        check(self.lib.msi_net_forward(self._h, ptr(psv), ptr(hi), ptr(lo), B, ptr(out), stream_ptr()),
              "msi_net_forward")
        return out if self.c_out_eng == self.c_out else out[..., :self.c_out]

    @property
    def can_fuse_rgba(self):
        return bool(self.lib.msi_net_can_fuse_rgba(self._h))

    def forward_rgba(self, psv=None, *, hi_lo=None, out=None):
        """msi_net_forward_rgba: the forward with the `blend_psv` RGBA assembly (msi.py:130-147) fused into the
        head's epilogue.  Returns rgba [B,H,W,L,4] (L = c_out / 2); the prediction is never written."""
        if psv is not None:
            B = psv.shape[0]
            psv = psv.contiguous()
        else:
            B = hi_lo[0].shape[0]
        L = self.c_out // 2
        if out is None:
            out = torch.empty((B, self.H, self.W, L, 4), dtype=torch.float32, device=self.device)
        hi, lo = hi_lo if hi_lo is not None else (None, None)
        check(self.lib.msi_net_forward_rgba(self._h, ptr(psv), ptr(hi), ptr(lo), B, ptr(out), stream_ptr()),
              "msi_net_forward_rgba")
        return out

    def read_activation(self, scope, B=1):
        from .nets import layer_channels, layer_geometry
        ch = layer_channels(self.c_in, self.c_out, self.ngf)[scope]
        h, w = layer_geometry(self.H, self.W)[scope]
        out = torch.empty((B, h, w, ch), dtype=torch.float32, device=self.device)
        check(self.lib.msi_net_read_activation(self._h, scope.encode(), B, ptr(out), stream_ptr()), "read_activation")
        return out

    def read_raw(self, scope, B=1):
        from .nets import layer_channels, layer_geometry
        ch = layer_channels(self.c_in, self.c_out, self.ngf)[scope]
        h, w = layer_geometry(self.H, self.W)[scope]
        out = torch.empty((B, h, w, ch), dtype=torch.float32, device=self.device)
        check(self.lib.msi_net_read_raw(self._h, scope.encode(), B, ptr(out), stream_ptr()), "read_raw")
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self.lib.msi_net_destroy(self._h)
                self._h = c_void_p()
        except Exception:
            pass


class MSIPipeline:
    """End-to-end frames on one GPU: (ref, src) ODS pair -> rendered ERP view + depth.

    Stages (all kernels of libmsi_b200.so, enqueued on one stream, graph-captured per batch size):
      K1 msi_psv_gather     -> PSV as the net's fp16 hi/lo operand, written straight into the net input; the sample
                               coordinates come from the per-rig table (ops.sweep_table) -- ``static_rig=False``
                               evaluates them every frame instead (msi_psv_build: same bits)
      K2 msi_net_forward_rgba -> RGBA layers [B,H,W,L,4] straight out of the head's epilogue (`blend_psv`, L = 32 / 64);
         otherwise msi_net_forward -> pred [B,H,W,n_pred] (pixel stride = the engine's padded head width) and
      K4 msi_rgba_assemble_strided -> RGBA layers
      K5 msi_render_composite -> rgb / depth (float32 + uint8)
    ``coord_net`` picks nets.msi_coord_train_net / nets.msi_train_net (FLAGS.coord_net, msi.py:120-127) and
    ``which_color_pred`` the head and assembly (msi.py:107-273), as MSI.infer_msi does.
    """

    def __init__(self, weights, H=320, W=640, num_planes=32, ngf=64, batch=1, device="cuda",
                 min_depth=1.0, max_depth=100.0, conv_impl="tcgen05", precision="fp16_fp8x",
                 img_dtype=torch.float32, use_graph=True, coord_net=True, which_color_pred="blend_psv",
                 static_rig=True, fuse_rgba=True):
        _lib.require_cuda()
        from .msi import MSI
        self.device = torch.device(device)
        self.H, self.W, self.P, self.B = H, W, num_planes, batch
        self.planes = MSI().inv_depths(min_depth, max_depth, num_planes)
        self.which_color_pred = which_color_pred
        self.n_pred = ops.color_pred_channels(which_color_pred, num_planes)
        self.net = NetEngine(weights, H, W, 6 * num_planes, self.n_pred, ngf, self.device, max_batch=batch,
                             conv_impl=conv_impl, precision=precision, variant="coord" if coord_net else "wrap")
        dev = self.device
        B = batch
        self.img_dtype = img_dtype
        self.ref = torch.empty((B, H, W, 3), dtype=img_dtype, device=dev)
        self.src = torch.empty((B, H, W, 3), dtype=img_dtype, device=dev)
        self.poses = torch.eye(4, device=dev).reshape(1, 1, 16).repeat(B, 2, 1).contiguous()
        self.baselines = torch.full((B,), 0.032, device=dev)
        # host mirror of the rig = the key of the cached coordinate table (no device read-back to look it up)
        self.static_rig = static_rig
        self._rig = (np.tile(np.eye(4, dtype=np.float32).reshape(1, 1, 16), (B, 2, 1)), np.full((B,), 0.032, np.float32))
        self._table = None
        self.tgt_pose_rt = torch.eye(4, device=dev).reshape(1, 16).repeat(B, 1).contiguous()
        self.tgt_pos = torch.zeros((B, 3), device=dev)
        self.depths = torch.tensor(self.planes, dtype=torch.float32, device=dev)
        if coord_net:
            self.hi, self.lo = self.net.input_buffers(B)   # the sweep kernel writes the net's operand in place
        else:                                              # wrap-padded input: dense operand, copied in by the forward
            self.hi = torch.zeros((B, H, W, self.net.in_c_stride), dtype=torch.float16, device=dev)
            self.lo = torch.zeros_like(self.hi)
        self.psv_scratch = ops.psv_scratch(B, H, W, dev)
        # `blend_psv` on the tensor-core back end: the head's epilogue assembles the RGBA layers (no `pred`, no K4 launch)
        self.fused_rgba = bool(fuse_rgba and which_color_pred == "blend_psv" and self.net.can_fuse_rgba)
        # the engine's head may be wider than n_pred (padded for the kernels' channel tiling): K4 reads it in place
        self.pred_buf = torch.empty((B, H, W, self.net.c_out_eng), dtype=torch.float32, device=dev)
        self.pred = self.pred_buf[..., :self.n_pred]
        self.rgba = torch.empty((B, H, W, num_planes, 4), dtype=torch.float32, device=dev)
        self.out = {
            "rgb": torch.empty((B, H, W, 3), dtype=torch.float32, device=dev),
            "depth": torch.empty((B, H, W, 3), dtype=torch.float32, device=dev),
            "rgb_u8": torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev),
            "depth_u8": torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev),
        }
        # pinned host staging for the end-to-end path
        np_dt = torch.uint8 if img_dtype == torch.uint8 else torch.float32
        self.h_ref = torch.empty((B, H, W, 3), dtype=np_dt).pin_memory()
        self.h_src = torch.empty((B, H, W, 3), dtype=np_dt).pin_memory()
        self.h_rgb_u8 = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
        self.h_depth_u8 = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
        self.use_graph = use_graph
        self._graph = None
        self.gather = None   # FrameGather: the render kernel also stores its uint8 view into every rank's gathered buffer
        self.launches_per_step = self.net.launches_per_forward + (3 if self.fused_rgba else 4)

    # -- device-resident step ---------------------------------------------------------------
    def _stages(self):
        """The four stage launchers of one step, in order, as (name, callable) pairs."""
        lib = _lib.load()
        B, H, W, P = self.B, self.H, self.W, self.P
        tb = ops.erp_tables(H, W, self.device)
        dt = _lib.IMG_U8 if self.img_dtype == torch.uint8 else _lib.IMG_F32

        def k1():
            if self.static_rig:
                t = self.sweep_table()
                check(lib.msi_psv_gather(ptr(self.ref), ptr(self.src), dt, 1, ptr(t.table), t.frames, B, H, W, P, None,
                                         ptr(self.hi), ptr(self.lo), self.net.in_c_stride, ptr(self.psv_scratch),
                                         self.psv_scratch.numel(), stream_ptr()), "msi_psv_gather")
                return
            check(lib.msi_psv_build(ptr(self.ref), ptr(self.src), dt, 1, ptr(self.poses), ptr(self.baselines),
                                    ptr(self.depths), *tb.ptrs(), B, H, W, P, None, ptr(self.hi), ptr(self.lo),
                                    self.net.in_c_stride, ptr(self.psv_scratch), self.psv_scratch.numel(), stream_ptr()),
                  "msi_psv_build")

        def k2():
            if self.fused_rgba:
                self.net.forward_rgba(hi_lo=(self.hi, self.lo), out=self.rgba)
            else:
                self.net.forward(hi_lo=(self.hi, self.lo), out=self.pred_buf)

        def k4():
            check(lib.msi_rgba_assemble_strided(ptr(self.pred_buf), self.n_pred, self.net.c_out_eng, None, ptr(self.hi),
                                                ptr(self.lo), self.net.in_c_stride, B, H, W, P,
                                                _lib.COLOR_MODES[self.which_color_pred], ptr(self.rgba), None, None, None,
                                                stream_ptr()), "msi_rgba_assemble_strided")

        def k5():
            g = self.gather
            if g is None:
                check(lib.msi_render_composite(ptr(self.rgba), ptr(self.tgt_pose_rt), ptr(self.tgt_pos), ptr(self.depths),
                                               *tb.ptrs(), B, H, W, P, ptr(self.out["rgb"]), ptr(self.out["depth"]),
                                               ptr(self.out["rgb_u8"]), ptr(self.out["depth_u8"]), stream_ptr()),
                      "msi_render_composite")
            else:   # fused with the output all-gather: peer / multicast stores from the render kernel
                check(lib.msi_render_composite_gather(
                    ptr(self.rgba), ptr(self.tgt_pose_rt), ptr(self.tgt_pos), ptr(self.depths), *tb.ptrs(), B, H, W, P,
                    ptr(self.out["rgb"]), ptr(self.out["depth"]), ptr(self.out["rgb_u8"]), ptr(self.out["depth_u8"]),
                    c_void_p(g.peers_dev), g.world, c_void_p(g.multicast_ptr) if g.multicast_ptr else None,
                    g.first_frame, stream_ptr()), "msi_render_composite_gather")

        if self.fused_rgba:
            return [("psv_build", k1), ("net", k2), ("render_composite", k5)]
        return [("psv_build", k1), ("net", k2), ("rgba_assemble", k4), ("render_composite", k5)]

    def sweep_table(self):
        """The cached sweep coordinates of the current rig (built on first use, shared between pipelines of the
        same rig through the ops-level cache)."""
        if self._table is None:
            self._table = ops.sweep_table(self._rig[0], self._rig[1], self.planes, self.H, self.W, self.device)
        return self._table

    def set_rig(self, poses=None, baselines=None):
        """Change the eye poses ([B,2,4,4] = pose_eye . ref_pose_inv, msi.py:1113-1127) and / or the ODS baselines
        ([B]).  Drops the coordinate table and the captured graph (the next step rebuilds both)."""
        B = self.B
        if poses is not None:
            p = np.ascontiguousarray(np.asarray(poses, np.float32).reshape(B, 2, 16))
            self.poses.copy_(torch.from_numpy(p))
            self._rig = (p, self._rig[1])
        if baselines is not None:
            b = np.ascontiguousarray(np.broadcast_to(np.asarray(baselines, np.float32).reshape(-1), (B,)))
            self.baselines.copy_(torch.from_numpy(b))
            self._rig = (self._rig[0], b)
        self._table = None
        self._graph = None

    def _enqueue(self):
        for name, fn in self._stages():
            with nvtx_range(name):
                fn()

    def stage_times(self, reps=5):
        """Milliseconds per launch of each stage, timed alone on the current stream: ``reps`` back-to-back
        launches between two CUDA events after one warm launch.  (bench.py: HBM roofline of K1 / K4 / K5.)"""
        out = {}
        for name, fn in self._stages():
            fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            e1.synchronize()
            out[name] = e0.elapsed_time(e1) / reps
        return out

    def attach_gather(self, gather):
        """Fuse the path's collective into the render kernel (see FrameGather).  Call before the first step()."""
        assert self._graph is None, "attach_gather() must precede the first graph-captured step()"
        assert gather.B == self.B and gather.H == self.H and gather.W == self.W
        self.gather = gather

    def step(self):
        """One pass of the hot path over the resident batch (inputs already in HBM)."""
        if self.static_rig:
            self.sweep_table()   # (outside any capture: building it synchronises)
        if not self.use_graph:
            self._enqueue()
            return
        if self._graph is None:
            # warm-up outside capture (lazy module load, cudaFuncSetAttribute), then capture
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._enqueue()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue()
            self._graph = g
        self._graph.replay()

    def set_inputs(self, ref, src, tgt_pos=None, baselines=None):
        ref, src = torch.as_tensor(ref), torch.as_tensor(src)
        if self.img_dtype == torch.uint8 and (ref.is_floating_point() or src.is_floating_point()):
            raise _lib.MsiError("set_inputs: this pipeline takes uint8 images; a float [0, 1] image would be truncated "
                                "to 0 / 1 by the cast (build the pipeline with img_dtype=torch.float32)")
        self.ref.copy_(ref.to(self.ref.dtype), non_blocking=True)
        self.src.copy_(src.to(self.src.dtype), non_blocking=True)
        if tgt_pos is not None:
            self.tgt_pos.copy_(torch.as_tensor(tgt_pos, dtype=torch.float32))
        if baselines is not None:
            self.set_rig(baselines=_host_f32(baselines))

    # -- end-to-end step: host buffers in, host buffers out --------------------------------
    def step_e2e(self, ref_host=None, src_host=None):
        """Host images -> pinned staging -> H2D -> hot path -> D2H of the rendered uint8 view and
        depth.  Returns (rgb_u8, depth_u8) pinned host tensors (valid after the sync inside)."""
        if ref_host is not None:
            self.h_ref.copy_(torch.as_tensor(ref_host))
            self.h_src.copy_(torch.as_tensor(src_host))
        self.ref.copy_(self.h_ref, non_blocking=True)
        self.src.copy_(self.h_src, non_blocking=True)
        self.step()
        self.h_rgb_u8.copy_(self.out["rgb_u8"], non_blocking=True)
        self.h_depth_u8.copy_(self.out["depth_u8"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.h_rgb_u8, self.h_depth_u8

    # -- pipelined end-to-end: copies of neighbouring frames overlap with compute ---------------
    def _init_streaming(self, depth=2):
        dev = self.device
        B, H, W = self.B, self.H, self.W
        dt = self.ref.dtype
        self._depth = depth
        self._copy_s = torch.cuda.Stream(device=dev)
        self._d2h_s = torch.cuda.Stream(device=dev)
        self._in_stage = [(torch.empty((B, H, W, 3), dtype=dt, device=dev), torch.empty((B, H, W, 3), dtype=dt, device=dev))
                          for _ in range(depth)]
        self._out_stage = [(torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev),
                            torch.empty((B, H, W, 3), dtype=torch.uint8, device=dev)) for _ in range(depth)]
        self._h_out = [(torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory(),
                        torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()) for _ in range(depth)]
        # pinned staging for pageable inputs, one pair per slot: the H2D copy out of it is asynchronous, so a
        # single pair would let the host overwrite batch i's pixels with batch i+1's while the DMA still reads them
        self._h_in = [(torch.empty((B, H, W, 3), dtype=dt).pin_memory(), torch.empty((B, H, W, 3), dtype=dt).pin_memory())
                      for _ in range(depth)]
        mk = lambda: [torch.cuda.Event() for _ in range(depth)]
        self._ev_h2d, self._ev_in_free, self._ev_done, self._ev_d2h = mk(), mk(), mk(), mk()
        self._submitted = 0
        self._collected = 0

    def submit(self, ref_host, src_host, after_compute=None):
        """Enqueue one batch: H2D on a copy stream, the hot path on the current stream, D2H of the
        uint8 view + depth on a third stream.  ``ref_host`` / ``src_host`` should be pinned (they are
        staged through pinned memory otherwise).  Up to ``depth`` batches are in flight; call
        ``collect()`` for every submit, in order.  ``after_compute`` (optional callable) runs on the
        compute stream right after the path -- the multi-GPU all-gather hooks in there."""
        if not hasattr(self, "_depth"):
            self._init_streaming()
        if self._submitted - self._collected >= self._depth:
            raise _lib.MsiError("submit(): pipeline full, call collect() first")
        k = self._submitted % self._depth
        cur = torch.cuda.current_stream(self.device)
        ref_host, src_host = torch.as_tensor(ref_host), torch.as_tensor(src_host)
        if not (ref_host.is_pinned() and src_host.is_pinned()):
            self._ev_h2d[k].synchronize()   # the previous H2D out of this slot's staging pair has finished
            h_ref, h_src = self._h_in[k]
            if not ref_host.is_pinned():
                h_ref.copy_(ref_host)
                ref_host = h_ref
            if not src_host.is_pinned():
                h_src.copy_(src_host)
                src_host = h_src
        in_ref, in_src = self._in_stage[k]
        with torch.cuda.stream(self._copy_s):
            self._copy_s.wait_event(self._ev_in_free[k])      # compute has consumed this staging slot
            in_ref.copy_(ref_host, non_blocking=True)
            in_src.copy_(src_host, non_blocking=True)
            self._ev_h2d[k].record(self._copy_s)
        cur.wait_event(self._ev_h2d[k])
        self.ref.copy_(in_ref, non_blocking=True)
        self.src.copy_(in_src, non_blocking=True)
        self._ev_in_free[k].record(cur)
        self.step()
        if after_compute is not None:
            after_compute()
        cur.wait_event(self._ev_d2h[k])                        # the previous D2H out of this slot is done
        o_rgb, o_dep = self._out_stage[k]
        o_rgb.copy_(self.out["rgb_u8"], non_blocking=True)
        o_dep.copy_(self.out["depth_u8"], non_blocking=True)
        self._ev_done[k].record(cur)
        with torch.cuda.stream(self._d2h_s):
            self._d2h_s.wait_event(self._ev_done[k])
            self._h_out[k][0].copy_(o_rgb, non_blocking=True)
            self._h_out[k][1].copy_(o_dep, non_blocking=True)
            self._ev_d2h[k].record(self._d2h_s)
        self._submitted += 1

    def collect(self):
        """Wait for the oldest in-flight batch; returns its (rgb_u8, depth_u8) pinned host tensors
        (valid until ``depth`` more batches have been submitted)."""
        if self._collected >= self._submitted:
            raise _lib.MsiError("collect(): nothing in flight")
        k = self._collected % self._depth
        self._ev_d2h[k].synchronize()
        self._collected += 1
        return self._h_out[k]

    @property
    def h2d_bytes_per_step(self):
        return self.h_ref.numel() * self.h_ref.element_size() * 2

    @property
    def d2h_bytes_per_step(self):
        return self.h_rgb_u8.numel() + self.h_depth_u8.numel()


class MSIFrameLanes:
    """Several frames in flight on one GPU: ``lanes`` independent ``MSIPipeline`` objects (own workspace,
    CUDA graph and stream) fed round-robin.

    The path is a chain of ~40 kernels per frame with two grid-wide dependencies per conv layer (the
    LayerNorm statistics, then the normalised activation), so a single frame leaves SMs idle at every
    kernel tail and ramp.  Frames are independent, so the next frame's kernels run on
    another stream and fill those bubbles: the bandwidth-bound kernels of one frame (LayerNorm, RGBA
    assembly, render) co-reside with the tensor-core kernels of the other.  Results are identical to
    the single-lane path (same kernels, same order per frame).  Measured on B200 (640x320x32, batch 1):
    669 frames/s one frame at a time, 738 / 752-767 / 776 with 2 / 3 / 4 lanes (single-CTA conv kernel); with the
    CTA-pair conv kernel 721 one frame at a time, 820 with 3 lanes.
    """

    def __init__(self, weights, *args, lanes=2, device="cuda", **kw):
        assert lanes >= 1
        self.device = torch.device(device)
        self.lanes = [MSIPipeline(weights, *args, device=device, **kw) for _ in range(lanes)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(lanes)]
        self._next_step = 0   # lane of the next step()
        self._next = 0        # lane of the next submit()
        self._oldest = 0      # lane of the next collect()
        self._fork_ev = torch.cuda.Event()

    def __len__(self):
        return len(self.lanes)

    def set_inputs(self, ref, src, tgt_pos=None, baselines=None):
        for p in self.lanes:
            p.set_inputs(ref, src, tgt_pos=tgt_pos, baselines=baselines)

    def fork(self):
        """The lane streams wait for everything enqueued so far on the current stream."""
        self._fork_ev.record(torch.cuda.current_stream(self.device))
        for s in self.streams:
            s.wait_event(self._fork_ev)

    def join(self):
        """The current stream waits for every lane."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)

    def step(self, after_compute=None):
        """One device-resident pass on the next lane (round-robin); returns that lane's pipeline."""
        k = self._next_step
        self._next_step = (k + 1) % len(self.lanes)
        with torch.cuda.stream(self.streams[k]):
            self.lanes[k].step()
            if after_compute is not None:
                after_compute(self.lanes[k])
        return self.lanes[k]

    def submit(self, ref_host, src_host, after_compute=None):
        """End-to-end: enqueue one batch on the next lane (see MSIPipeline.submit); collect() in order."""
        k = self._next
        self._next = (k + 1) % len(self.lanes)
        lane = self.lanes[k]
        with torch.cuda.stream(self.streams[k]):
            lane.submit(ref_host, src_host,
                        after_compute=(lambda: after_compute(lane)) if after_compute is not None else None)

    def collect(self):
        k = self._oldest
        self._oldest = (k + 1) % len(self.lanes)
        return self.lanes[k].collect()

    @property
    def in_flight_capacity(self):
        return sum(getattr(p, "_depth", 2) for p in self.lanes)


class FrameGather:
    """The gathered output buffer ``frames [world * B, H, W, 3]`` uint8 in symmetric memory (every rank
    maps every rank's copy), for the fused form of the path's only collective: the render
    kernel of rank r stores its B frames at index r * B of EVERY rank's buffer -- one ``multimem.st`` per
    word through the NVSwitch multicast address when the fabric offers one, else one peer store per rank
    -- so that no all-gather kernel (and no per-step rendezvous of the ranks) follows it.  Readers call
    ``barrier()`` before they look at ``frames`` and again before the producers may overwrite it.

    ``slots`` > 1: ONE symmetric allocation (one rendezvous, one multicast object) carved into that many
    independent gathered buffers, one per frame lane (``slot(k)``); a lane's render kernel addresses its
    slot through ``first_frame``.

    mode: "auto" (multicast if available, else peer stores), "peer", "multicast".  Raises if symmetric
    memory cannot be set up (the caller then falls back to ``all_gather_frames`` = NCCL)."""

    def __init__(self, B, H, W, device, group=None, mode="auto", slots=1):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.B, self.H, self.W, self.slots = B, H, W, slots
        assert (H * W * 3) % 4 == 0
        per_slot = self.world * B * H * W * 3
        self.buf = symm_mem.empty(slots * per_slot, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.peers_dev = int(self.hdl.buffer_ptrs_dev)
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        if mode == "multicast" and not mc:
            raise _lib.MsiError("FrameGather: no multicast address on this fabric")
        self.multicast_ptr = mc if mode in ("auto", "multicast") else 0
        self.mode = "multimem.st (NVSwitch multicast)" if self.multicast_ptr else "peer stores"
        self.all_frames = self.buf.view(slots, self.world * B, H, W, 3)
        self.frames = self.all_frames[0]
        self.first_frame = self.rank * B
        torch.cuda.synchronize(device)
        self.barrier()

    def slot(self, k):
        """The k-th gathered buffer of this allocation, with the attributes a pipeline's render launch reads."""
        if self.slots == 1 and k == 0:
            return self
        assert 0 <= k < self.slots
        v = object.__new__(FrameGather)
        v.__dict__.update(self.__dict__)
        v.frames = self.all_frames[k]
        v.first_frame = k * self.world * self.B + self.rank * self.B   # frame index inside the whole allocation
        v.slots = 1
        return v

    def barrier(self):
        """All ranks' stores issued before the barrier (on the current stream) are visible after it."""
        self.hdl.barrier(channel=0)


def shard_frames(num_frames: int, rank: int, world_size: int):
    """Frame range [lo, hi) of ``rank`` (rank r takes frames [r*B/N, (r+1)*B/N))."""
    per = -(-num_frames // world_size)
    lo = min(rank * per, num_frames)
    return lo, min(lo + per, num_frames)


def all_gather_frames(local: torch.Tensor, world_size: int, group=None) -> torch.Tensor:
    """The path's only collective: gather the rendered frames of every rank, rank-major.
    local: [b,H,W,3] (same b on every rank).  Works on NCCL (GPU) and gloo (CPU tests)."""
    import torch.distributed as dist
    if world_size == 1:
        return local
    out = torch.empty((world_size * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def profile_net_layers(net: NetEngine, hi_lo, out, reps: int = 3, rgba=None, cold_l2: bool = False):
    """Per-layer conv / LayerNorm milliseconds (CUDA events on the launching stream, average of
    ``reps`` forwards) and algorithmic FLOPs per layer.  Used by bench.py for the roofline.
    ``rgba``: run the head with the fused RGBA assembly (as the pipeline does); ``cold_l2``: overwrite a 256 MB
    buffer before every timed launch (each launch timed on its own) instead of timing back-to-back repeats."""
    lib = net.lib
    n = int(lib.msi_net_num_layers(net._h))
    conv = (ctypes.c_float * n)()
    ln = (ctypes.c_float * n)()
    B = hi_lo[0].shape[0]
    acc_c, acc_l = np.zeros(n), np.zeros(n)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=net.device) if cold_l2 else None
    for _ in range(reps):
        check(lib.msi_net_forward_profiled_flush(net._h, None, ptr(hi_lo[0]), ptr(hi_lo[1]), B, ptr(out), stream_ptr(),
                                                 conv, ln, ptr(flush), flush.numel() if flush is not None else 0,
                                                 ptr(rgba)), "msi_net_forward_profiled_flush")
        acc_c += np.frombuffer(conv, dtype=np.float32)
        acc_l += np.frombuffer(ln, dtype=np.float32)
    scopes = [lib.msi_net_layer_scope(net._h, i).decode() for i in range(n)]
    flops = np.array([lib.msi_net_layer_flops(net._h, i) for i in range(n)]) * B
    return scopes, acc_c / reps, acc_l / reps, flops
