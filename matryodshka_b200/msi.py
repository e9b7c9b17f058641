"""Host-side mirror of the reference's ``matryodshka/msi.py`` -- class ``MSI``, inference half.

Same method names, argument order and dict keys as the reference (msi.py:40-52, :276-289,
:384-452, :1094-1217) so that a driver written against ``matryodshka.msi.MSI`` reads the same.
What the reference takes from TF graph-global state is explicit here:

* the FLAGS it reads deep inside (``input_type``, ``operation``, ``coord_net``, ``ngf`` ...;
  msi.py:70,81,95-105,120-127) are a frozen ``MSIConfig``;
* the hidden graph tensors ``ref_pose_inv:0`` / ``jitter_pose_inv:0`` (msi.py:1113-1119) are the
  keyword arguments ``ref_pose_inv`` / ``jitter_pose_inv``;
* the TF variables under scope ``net/`` are a ``weights`` dict keyed by the checkpoint names.

Tensors are torch CUDA tensors, NHWC float32 (images in [0,1] at the API).  All arithmetic runs
in the sm_100a kernels behind include/msi_b200.h; there is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from . import torch_ops  # noqa: F401  (registers torch.ops.msi.*)
from .runtime import NetEngine, nvtx_range


@dataclass(frozen=True)
class MSIConfig:
    """The reference's flags for this path, same names (test.py:39-83, loader.py:30-42) and the same defaults
    EXCEPT ``coord_net``: the reference defaults to False (test.py:52) while its released model and BASELINE's
    configs are the coord net (``--coord_net``), which is the default here.  ``MSI`` checks the choice against the
    shape of the conv1_1 weights it is given and refuses a mismatch."""
    height: int = 320
    width: int = 640
    num_psv_planes: int = 32
    num_msi_planes: int = 32
    min_depth: float = 1.0
    max_depth: float = 100.0
    ngf: int = 64
    which_color_pred: str = "blend_psv"
    coord_net: bool = True
    input_type: str = "ODS"
    operation: str = "train"
    transform_inverse_reg: bool = False
    jitter: bool = False
    net_only: bool = False
    supervision: str = "tgt"
    batch_size: int = 1
    # back-end selection (ours)
    conv_impl: str = "tcgen05"
    precision: str = "fp16_fp8x"


def _matmul44_f32(a, b):
    """[N,4,4] x [N,4,4] in float32 with the k-sum evaluated left to right."""
    a = np.asarray(a, np.float32).reshape(-1, 4, 4)
    b = np.asarray(b, np.float32).reshape(-1, 4, 4)
    out = np.zeros_like(a)
    for i in range(4):
        for j in range(4):
            acc = a[:, i, 0] * b[:, 0, j]
            for k in range(1, 4):
                acc = acc + a[:, i, k] * b[:, k, j]
            out[:, i, j] = acc
    return out


def _host(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def _dev(x, device, shape=None):
    """Host array / tensor -> contiguous float32 tensor on ``device`` (the custom ops take tensors only)."""
    return ops._dev_f32(x, device, shape)


class MSI(object):
    """Multi-sphere-image inference (reference class ``MSI``, msi.py:33)."""

    def __init__(self, weights=None, config: Optional[MSIConfig] = None, device="cuda"):
        self.config = config or MSIConfig()
        self.weights = weights
        self.device = torch.device(device)
        self._engines = {}

    def load_weights(self, weights):
        """``MSI()`` as the reference constructs it (no arguments), weights attached afterwards: a dict keyed by the
        TF variable names, or the prefix of a TensorFlow checkpoint (what ``saver.restore`` takes, test.py:192-202)."""
        if isinstance(weights, (str, bytes)):
            from .tf_checkpoint import load_checkpoint
            weights = {k: v for k, v in load_checkpoint(weights).items() if k.startswith("net/")}
        self.weights = weights
        self._engines = {}
        return self

    # ---- msi.py:1196-1217 ------------------------------------------------------------------
    def inv_depths(self, start_depth, end_depth, num_depths):
        """Depths uniform in inverse depth, sorted far -> near (python floats)."""
        inv_start_depth = 1.0 / start_depth
        inv_end_depth = 1.0 / end_depth
        depths = [start_depth, end_depth]
        for i in range(1, num_depths - 1):
            fraction = float(i) / float(num_depths - 1)
            inv_depth = inv_start_depth + (inv_end_depth - inv_start_depth) * fraction
            depths.append(1.0 / inv_depth)
        depths = sorted(depths)
        return depths[::-1]

    # ---- msi.py:1163-1194 --------------------------------------------------------------------
    def preprocess_image(self, image):
        """float [0,1] or uint8 [0,255] -> float32 [-1,1].  (On the fused path this happens inside
        msi_psv_build; this stand-alone form is API glue on a [B,H,W,3] tensor.)"""
        if image.dtype == torch.uint8:
            image = image.to(torch.float32) * np.float32(1.0 / 255)
        return image.to(torch.float32) * 2 - 1

    def deprocess_image(self, image):
        """float32 [-1,1] -> uint8, tf.image.convert_image_dtype semantics (x*255.5 truncated)."""
        image = (image + 1.0) / 2.0
        return (image * 255.5).to(torch.int32).to(torch.uint8)

    def deprocess_depth_image(self, image):
        return (image * 255.5).to(torch.int32).to(torch.uint8)

    # ---- msi.py:1094-1130 ----------------------------------------------------------------------
    def _sweep_poses(self, ref_pose, src_pose, ref_pose_inv=None, jitter_pose_inv=None):
        ref_pose = _host(ref_pose).astype(np.float32).reshape(-1, 4, 4)
        src_pose = _host(src_pose).astype(np.float32).reshape(-1, 4, 4)
        if ref_pose_inv is None:
            ref_pose_inv = np.linalg.inv(ref_pose.astype(np.float64)).astype(np.float32)
        ref_pose_inv = _host(ref_pose_inv).astype(np.float32).reshape(-1, 4, 4)
        if jitter_pose_inv is not None:
            ref_pose_inv = _matmul44_f32(ref_pose_inv, _host(jitter_pose_inv))
        eye0 = _matmul44_f32(ref_pose, ref_pose_inv)
        eye1 = _matmul44_f32(src_pose, ref_pose_inv)
        return np.stack([eye0, eye1], axis=1)  # [B,2,4,4]

    def format_network_input(self, ref_image, src_image, ref_pose, src_pose, planes, intrinsics,
                             ref_pose_inv=None, jitter_pose_inv=None):
        """Double plane-sweep volume [B,H,W,6P] from preprocessed ([-1,1]) images; channel =
        eye*3P + p*3 + rgb, eye 0 = ref (order +1), eye 1 = src (order -1)."""
        poses = self._sweep_poses(ref_pose, src_pose, ref_pose_inv, jitter_pose_inv)
        baselines = _host(intrinsics).astype(np.float32).reshape(-1, 3, 3)[:, 0, 0]
        dev = ref_image.device
        return torch.ops.msi.psv_build(ref_image, src_image, _dev(poses, dev), _dev(baselines, dev),
                                       _dev(list(planes), dev), False)

    def sweep_src(self, image, order, depths, pose, intrinsics):
        from .geometry import projector as pj
        return pj.ods_sphere_sweep(image, order, depths, pose, intrinsics)

    # ---- msi.py:40-289 ---------------------------------------------------------------------------
    def _engine(self, H, W, c_in, c_out, ngf, max_batch):
        key = (H, W, c_in, c_out, ngf, max_batch)
        eng = self._engines.get(key)
        if eng is None:   # an engine built for a larger batch serves a smaller one
            for k2, e2 in self._engines.items():
                if k2[:5] == key[:5] and k2[5] >= max_batch:
                    eng = e2
                    break
        if eng is None:
            if self.weights is None:
                raise _lib.MsiError("MSI.infer_msi needs weights (MSI(weights=...) or MSI().load_weights(...))")
            w0 = self.weights.get("net/conv1_1/weights")
            if w0 is not None and tuple(w0.shape)[2] != c_in + (1 if self.config.coord_net else 0):
                raise _lib.MsiError(
                    "MSIConfig.coord_net=%s but net/conv1_1/weights has %d input channels for a %d-channel PSV (%s): "
                    "set MSIConfig(coord_net=%s)" % (self.config.coord_net, tuple(w0.shape)[2], c_in,
                                                     "coord net" if tuple(w0.shape)[2] == c_in + 1 else "msi_train_net",
                                                     tuple(w0.shape)[2] == c_in + 1))
            # FLAGS.coord_net picks nets.msi_coord_train_net or nets.msi_train_net (msi.py:120-127)
            eng = NetEngine(self.weights, H, W, c_in, c_out, ngf, self.device, max_batch=max_batch,
                            conv_impl=self.config.conv_impl, precision=self.config.precision,
                            variant="coord" if self.config.coord_net else "wrap")
            self._engines[key] = eng
        return eng

    def infer_msi(self, raw_src_image, raw_ref_image, raw_hres_src_image, raw_hres_ref_image,
                  ref_pose, src_pose, intrinsics, which_color_pred, num_msi_planes, psv_planes,
                  extra_outputs='', ngf=64, ref_pose_inv=None, jitter_pose_inv=None):
        """Same signature as the reference (argument order src, ref).  Returns (pred dict, net_input)
        with pred['rgba_layers'] [B,H,W,L,4] always and 'blend_weights' / 'alphas' / 'psv' when the
        substring is in ``extra_outputs`` (msi.py:276-288)."""
        cfg = self.config
        if which_color_pred not in ('blend_psv', 'blend_bg', 'blend_bg_psv', 'alpha_only'):
            raise NotImplementedError("which_color_pred=%r (msi.py:107-116 has blend_psv, blend_bg, blend_bg_psv, "
                                      "alpha_only)" % (which_color_pred,))
        if cfg.input_type != 'ODS' or cfg.operation != 'train':
            raise NotImplementedError("only input_type=ODS, operation=train is built")
        B, H, W, _ = raw_src_image.shape
        P = len(psv_planes)
        if P != num_msi_planes:
            raise ValueError("blend_psv needs len(psv_planes) == num_msi_planes")
        poses = self._sweep_poses(ref_pose, src_pose, ref_pose_inv, jitter_pose_inv)
        baselines = _host(intrinsics).astype(np.float32).reshape(-1, 3, 3)[:, 0, 0]
        eng = self._engine(H, W, 6 * P, ops.color_pred_channels(which_color_pred, num_msi_planes), ngf, B)
        if cfg.coord_net:
            hi, lo = eng.input_buffers(B)   # the sweep kernel writes the net's operand in place
        else:                               # wrap-padded input: dense operand, copied in by the forward
            hi = torch.zeros((B, H, W, eng.in_c_stride), dtype=torch.float16, device=self.device)
            lo = torch.zeros_like(hi)
        # preprocessing (msi.py:73-75) is fused into the sweep kernel
        # a static rig (every frame of the reference's data: identity eye poses, one baseline) gathers from the
        # cached coordinate table; the jittered sweep evaluates the chain per call.  Same bits either way.
        with nvtx_range("psv_build"):
            net_input = ops.psv_build(raw_ref_image, raw_src_image, poses, baselines, list(psv_planes),
                                      preprocess=True, want_f32=True, hi_lo=(hi, lo), c_stride=eng.in_c_stride,
                                      cache_coords=jitter_pose_inv is None)
        if cfg.net_only:
            eng.forward(hi_lo=(hi, lo))
            return None
        with nvtx_range("net"):
            msi_pred = eng.forward(hi_lo=(hi, lo))
        want_w = ('blend_weights' in extra_outputs) or ('alpha' in extra_outputs)
        bgw = None
        if which_color_pred == 'blend_psv':
            rgba, bw, al = ops.rgba_assemble(msi_pred, net_input, want_weights=want_w)
        else:
            rgba, bw, al, bgw = ops.rgba_assemble_ex(msi_pred, net_input, which_color_pred, num_msi_planes,
                                                     want_weights=want_w)
        pred = {'rgba_layers': rgba}
        if 'blend_weights' in extra_outputs and 'blend' in which_color_pred:
            pred['blend_weights'] = bw
            if bgw is not None:   # msi.py:281-282 (the reference names it for every 'bg' scheme; only blend_bg_psv defines it)
                pred['bg_blend_weights'] = bgw
        if 'alpha' in extra_outputs:
            pred['alphas'] = al
        if 'psv' in extra_outputs:
            pred['psv'] = net_input
        return pred, net_input

    # ---- msi.py:384-452 ------------------------------------------------------------------------
    def msi_render_equirect_view(self, rgba_layers, tgt_pose_rt, tgt_pos, planes, intrinsics=None):
        """Rendered view [B,H,W,3] in [-1,1]."""
        return self._render(rgba_layers, tgt_pose_rt, tgt_pos, planes)[0]

    def msi_render_equirect_depth(self, rgba_layers, tgt_pose_rt, tgt_pos, planes, intrinsics=None):
        """Composited normalised layer index [B,H,W,3] in [0,1)."""
        return self._render(rgba_layers, tgt_pose_rt, tgt_pos, planes)[1]

    @staticmethod
    def _render(rgba_layers, tgt_pose_rt, tgt_pos, planes):
        """torch.ops.msi.render_composite: (rgb, depth, rgb_u8, depth_u8) from ONE reprojection."""
        dev = rgba_layers.device
        B = rgba_layers.shape[0]
        return torch.ops.msi.render_composite(rgba_layers.contiguous(), _dev(tgt_pose_rt, dev, (B, 16)),
                                              _dev(tgt_pos, dev, (B, 3)), _dev(list(planes), dev))

    def msi_render_equirect(self, rgba_layers, tgt_pose_rt, tgt_pos, planes, intrinsics=None):
        """Ours: colour and depth (float32 + deprocessed uint8) from ONE reprojection."""
        rgb, depth, rgb_u8, depth_u8 = self._render(rgba_layers, tgt_pose_rt, tgt_pos, planes)
        return {"rgb": rgb, "depth": depth, "rgb_u8": rgb_u8, "depth_u8": depth_u8}

    def msi_render_equirect_view_single(self, rgba_layers, tgt_pose_rt, tgt_pos, planes, intrinsics=None):
        """Reprojected layers without compositing [L,B,H,W,4] (msi.py:431-452)."""
        planes = planes.tolist() if torch.is_tensor(planes) else list(planes)
        dev = rgba_layers.device
        B = rgba_layers.shape[0]
        return torch.ops.msi.project_layers(rgba_layers.contiguous(), _dev(tgt_pose_rt, dev, (B, 16)),
                                            _dev(tgt_pos, dev, (B, 3)), _dev(planes, dev))

    msi_render_equirect_depth_single = msi_render_equirect_view_single

    def msi_render_perspective_view(self, rgba_layers, tgt_pose_rt, tgt_pos, planes, intrinsics=None,
                                    viewing_window=3, psp_height=270, psp_width=480):
        """msi.py:475-500: pinhole view of the MSI.  As in the reference, ``tgt_pose_rt`` is accepted and
        replaced by the viewing-window rotation (projector.py:80-85) and the intrinsics are the
        hard-coded ones of spherical.py:385-387."""
        return ops.render_perspective(rgba_layers, _host(tgt_pos).reshape(-1, 3), list(planes),
                                      viewing_window=viewing_window, psp_height=psp_height, psp_width=psp_width)

    def msi_render_ods_view(self, rgba_layers, order, jitter_pose, tgt_pos, planes, intrinsics):
        """msi.py:502-525: the MSI seen from one ODS eye (order +1 = left / ref, -1 = right / src)
        under ``jitter_pose`` [B,4,4].  As in the reference, ``tgt_pos`` is not used by the ODS ray
        generator (spherical.intersect_ods) and the baseline is intrinsics[0][0][0] for every frame."""
        B = rgba_layers.shape[0]
        base = float(_host(intrinsics).astype(np.float32).reshape(-1, 3, 3)[0, 0, 0])
        return ops.render_ods(rgba_layers, jitter_pose, order, [base] * B, list(planes))
