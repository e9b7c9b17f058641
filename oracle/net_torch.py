"""Oracle (test infrastructure): torch-CPU restatement of the reference conv net.

Follows /root/reference matryodshka/nets.py: msi_coord_train_net (:471-515) with
coord_conv2d (:267-270) and add_sph_coords (:260-265); msi_train_net (:387-441)
with wrap_pad (:288-295).  The arithmetic of tf.contrib.slim [TF-1.14] is
restated explicitly:

* ``slim.conv2d`` with a normalizer has NO bias; with ``normalizer_fn=None`` it
  has a bias.  Weights are HWIO ``[kh, kw, Cin, Cout]``.
* SAME padding: ``pad_total = max((ceil(n / s) - 1) * s + k_eff - n, 0)`` with
  ``k_eff = (k - 1) * rate + 1``; ``before = pad_total // 2``.
* ``slim.conv2d_transpose`` 4x4 stride 2 SAME == the gradient of a SAME stride-2
  conv: ``conv_transpose2d(k=4, s=2, p=1)`` with
  ``w_pt[ci, co, kh, kw] = w_tf[kh, kw, co, ci]`` (TF layout ``[kh, kw, Cout, Cin]``).
* ``slim.layer_norm``: moments over axes 1..3 (two-pass variance), then
  ``tf.nn.batch_normalization`` with eps 1e-12:
  ``inv = rsqrt(var + eps) * gamma; y = x * inv + (beta - mean * inv)``;
  gamma, beta of shape [C].

Tensors are NHWC at the API, as in the reference.  Variable names follow the TF
checkpoint: ``net/<scope>/weights``, ``net/<scope>/LayerNorm/{beta,gamma}``,
``net/color_pred/biases``.

PINNED (1e-5) to the reference's own nets.py run over a restated slim op layer (tests/golden/reference_run.npz) -- see oracle/__init__.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# (scope, kind, cout multiplier of ngf, kernel, stride, rate) -- nets.py:486-507
LAYERS = [
    ("conv1_1", "conv", 1, 3, 1, 1),
    ("conv1_2", "conv", 2, 3, 2, 1),
    ("conv2_1", "conv", 2, 3, 1, 1),
    ("conv2_2", "conv", 4, 3, 2, 1),
    ("conv3_1", "conv", 4, 3, 1, 1),
    ("conv3_2", "conv", 4, 3, 1, 1),
    ("conv3_3", "conv", 8, 3, 2, 1),
    ("conv4_1", "conv", 8, 3, 1, 2),
    ("conv4_2", "conv", 8, 3, 1, 2),
    ("conv4_3", "conv", 8, 3, 1, 2),
    ("conv6_1", "deconv", 4, 4, 2, 1),
    ("conv6_2", "conv", 4, 3, 1, 1),
    ("conv6_3", "conv", 4, 3, 1, 1),
    ("conv7_1", "deconv", 2, 4, 2, 1),
    ("conv7_2", "conv", 2, 3, 1, 1),
    ("conv8_1", "deconv", 1, 4, 2, 1),
    ("conv8_2", "conv", 1, 3, 1, 1),
]


def same_pad(n, k, s, rate=1):
    """[TF-1.14] SAME padding (before, after) for one spatial dim."""
    k_eff = (k - 1) * rate + 1
    out = -(-n // s)
    total = max((out - 1) * s + k_eff - n, 0)
    return total // 2, total - total // 2


def sph_coord_rows(height):
    """nets.py:260-263: ``|sin(linspace(-pi/2, pi/2, H))|`` computed in float64
    NumPy, cast to float32.  (The ``+ input[..., :1] / sys.float_info.max`` term
    is ``x / 1.8e308`` = 0 in float32.)"""
    return np.abs(np.sin(np.linspace(-np.pi / 2.0, np.pi / 2.0, height))).astype(np.float32)


def add_sph_coords(x):
    """nets.py:260-265 on an NHWC tensor."""
    B, H, W, _ = x.shape
    coord = torch.from_numpy(sph_coord_rows(H)).to(x.dtype).view(1, H, 1, 1).expand(B, H, W, 1)
    return torch.cat([x, coord], dim=3)


def conv2d_same(x, w_hwio, stride=1, rate=1, bias=None, padding="SAME"):
    """slim.conv2d forward on NHWC input with HWIO weights [TF-1.14]."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    xn = x.permute(0, 3, 1, 2)
    if padding == "SAME":
        pt, pb = same_pad(x.shape[1], kh, stride, rate)
        pl, pr = same_pad(x.shape[2], kw, stride, rate)
        xn = F.pad(xn, (pl, pr, pt, pb))
    w = w_hwio.permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(xn, w, bias=bias, stride=stride, dilation=rate)
    return y.permute(0, 2, 3, 1).contiguous()


def conv2d_transpose_same(x, w_hwoi, stride=2):
    """slim.conv2d_transpose 4x4 stride-2 SAME on NHWC input; weights in the TF
    layout [kh, kw, Cout, Cin] [TF-1.14]."""
    assert w_hwoi.shape[0] == 4 and w_hwoi.shape[1] == 4 and stride == 2
    xn = x.permute(0, 3, 1, 2)
    w = w_hwoi.permute(3, 2, 0, 1).contiguous()  # [Cin, Cout, kh, kw]
    y = F.conv_transpose2d(xn, w, stride=2, padding=1)
    return y.permute(0, 2, 3, 1).contiguous()


def layer_norm_relu(x, gamma, beta, eps=1e-12, relu=True):
    """slim.layer_norm (begin_norm_axis=1, begin_params_axis=-1) + ReLU [TF-1.14]."""
    mean = x.mean(dim=(1, 2, 3), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    inv = torch.rsqrt(var + eps) * gamma.view(1, 1, 1, -1)
    y = x * inv + (beta.view(1, 1, 1, -1) - mean * inv)
    return F.relu(y) if relu else y


def wrap_pad(x, left_pad, right_pad):
    """nets.py:288-295: circular pad in width, zero pad in height (NHWC)."""
    left = x[:, :, -left_pad:, :]
    right = x[:, :, :right_pad, :]
    x = torch.cat([left, x, right], dim=2)
    return F.pad(x, (0, 0, 0, 0, left_pad, right_pad))


def layer_shapes(num_inputs, num_outputs, ngf=64, coord=True):
    """Weight shapes in TF layout keyed by TF variable name."""
    shapes = {}
    cin = num_inputs
    feats = {}
    for scope, kind, mult, k, s, r in LAYERS:
        cout = ngf * mult
        if kind == "conv":
            shapes[f"net/{scope}/weights"] = (k, k, cin + (1 if coord else 0), cout)
        else:
            if scope == "conv6_1":
                cin = feats["conv4_3"] + feats["conv3_3"]
            elif scope == "conv7_1":
                cin = feats["conv6_3"] + feats["conv2_2"]
            elif scope == "conv8_1":
                cin = feats["conv7_2"] + feats["conv1_2"]
            shapes[f"net/{scope}/weights"] = (k, k, cout, cin)
        shapes[f"net/{scope}/LayerNorm/gamma"] = (cout,)
        shapes[f"net/{scope}/LayerNorm/beta"] = (cout,)
        feats[scope] = cout
        cin = cout
    shapes["net/color_pred/weights"] = (1, 1, cin, num_outputs)
    shapes["net/color_pred/biases"] = (num_outputs,)
    return shapes


def _t(weights, name, dtype):
    w = weights[name]
    if not torch.is_tensor(w):
        w = torch.from_numpy(np.asarray(w))
    return w.to(dtype)


def msi_coord_train_net(inputs, num_outputs, weights, ngf=64, dtype=torch.float32,
                        quant=None, return_feats=False):
    """nets.py:471-515.  inputs: NHWC torch tensor; weights: dict keyed by TF
    variable names.  ``quant`` (optional callable) is applied to every conv
    operand (activations and weights) -- used only to budget reduced-precision
    operand error against the float32 result; None for the oracle proper."""
    q = quant if quant is not None else (lambda t: t)
    x = inputs.to(dtype)
    feats = {}

    def cconv(x, scope, stride=1, rate=1):
        w = _t(weights, f"net/{scope}/weights", dtype)
        xin = add_sph_coords(x)
        y = conv2d_same(q(xin), q(w), stride=stride, rate=rate)
        y = layer_norm_relu(y, _t(weights, f"net/{scope}/LayerNorm/gamma", dtype),
                            _t(weights, f"net/{scope}/LayerNorm/beta", dtype))
        feats[scope] = y
        return y

    def deconv(x, scope):
        w = _t(weights, f"net/{scope}/weights", dtype)
        y = conv2d_transpose_same(q(x), q(w))
        y = layer_norm_relu(y, _t(weights, f"net/{scope}/LayerNorm/gamma", dtype),
                            _t(weights, f"net/{scope}/LayerNorm/beta", dtype))
        feats[scope] = y
        return y

    c11 = cconv(x, "conv1_1")
    c12 = cconv(c11, "conv1_2", stride=2)
    c21 = cconv(c12, "conv2_1")
    c22 = cconv(c21, "conv2_2", stride=2)
    c31 = cconv(c22, "conv3_1")
    c32 = cconv(c31, "conv3_2")
    c33 = cconv(c32, "conv3_3", stride=2)
    c41 = cconv(c33, "conv4_1", rate=2)
    c42 = cconv(c41, "conv4_2", rate=2)
    c43 = cconv(c42, "conv4_3", rate=2)
    c61 = deconv(torch.cat([c43, c33], dim=3), "conv6_1")
    c62 = cconv(c61, "conv6_2")
    c63 = cconv(c62, "conv6_3")
    c71 = deconv(torch.cat([c63, c22], dim=3), "conv7_1")
    c72 = cconv(c71, "conv7_2")
    c81 = deconv(torch.cat([c72, c12], dim=3), "conv8_1")
    c82 = cconv(c81, "conv8_2")
    w = _t(weights, "net/color_pred/weights", dtype)
    b = _t(weights, "net/color_pred/biases", dtype)
    pred = torch.tanh(conv2d_same(q(c82), q(w), bias=b))
    if return_feats:
        return pred, feats
    return pred


def msi_train_net(inputs, num_outputs, weights, ngf=64, dtype=torch.float32, return_feats=False):
    """nets.py:387-441: the non-coord variant -- circular-x / zero-y padding
    (wrap_pad) + VALID convs; deconvs on wrap_pad(skip, 2, 2) cropped [5:-5].  Note the order in the
    reference (nets.py:431-436): slim.layer_norm runs inside conv2d_transpose on the FULL
    (2H+10) x (2W+10) output and the crop comes after it, so the statistics include the border."""
    x = inputs.to(dtype)

    def conv(x, scope, stride=1, rate=1):
        w = _t(weights, f"net/{scope}/weights", dtype)
        p = rate
        y = conv2d_same(wrap_pad(x, p, p), w, stride=stride, rate=rate, padding="VALID")
        return layer_norm_relu(y, _t(weights, f"net/{scope}/LayerNorm/gamma", dtype),
                               _t(weights, f"net/{scope}/LayerNorm/beta", dtype))

    def deconv(x, scope):
        w = _t(weights, f"net/{scope}/weights", dtype)
        xn = wrap_pad(x, 2, 2).permute(0, 3, 1, 2)
        wp = w.permute(3, 2, 0, 1).contiguous()
        y = F.conv_transpose2d(xn, wp, stride=2, padding=0).permute(0, 2, 3, 1).contiguous()  # VALID
        y = layer_norm_relu(y, _t(weights, f"net/{scope}/LayerNorm/gamma", dtype),
                            _t(weights, f"net/{scope}/LayerNorm/beta", dtype))
        return y[:, 5:-5, 5:-5, :]

    c11 = conv(x, "conv1_1")
    c12 = conv(c11, "conv1_2", stride=2)
    c21 = conv(c12, "conv2_1")
    c22 = conv(c21, "conv2_2", stride=2)
    c31 = conv(c22, "conv3_1")
    c32 = conv(c31, "conv3_2")
    c33 = conv(c32, "conv3_3", stride=2)
    c41 = conv(c33, "conv4_1", rate=2)
    c42 = conv(c41, "conv4_2", rate=2)
    c43 = conv(c42, "conv4_3", rate=2)
    c61 = deconv(torch.cat([c43, c33], dim=3), "conv6_1")
    c62 = conv(c61, "conv6_2")
    c63 = conv(c62, "conv6_3")
    c71 = deconv(torch.cat([c63, c22], dim=3), "conv7_1")
    c72 = conv(c71, "conv7_2")
    c81 = deconv(torch.cat([c72, c12], dim=3), "conv8_1")
    c82 = conv(c81, "conv8_2")
    w = _t(weights, "net/color_pred/weights", dtype)
    b = _t(weights, "net/color_pred/biases", dtype)
    pred = torch.tanh(conv2d_same(c82, w, bias=b))
    if return_feats:
        return pred, {"conv1_1": c11, "conv1_2": c12, "conv2_1": c21, "conv2_2": c22, "conv3_1": c31, "conv3_2": c32,
                      "conv3_3": c33, "conv4_1": c41, "conv4_2": c42, "conv4_3": c43, "conv6_1": c61, "conv6_2": c62,
                      "conv6_3": c63, "conv7_1": c71, "conv7_2": c72, "conv8_1": c81, "conv8_2": c82}
    return pred


def net_gflop(H, W, num_inputs, num_outputs, ngf=64, coord=True):
    """FLOPs of one forward (2*MACs), formula of SURVEY.md 8(a) a10 table."""
    shapes = layer_shapes(num_inputs, num_outputs, ngf, coord)
    h, w = H, W
    total = 0.0
    for scope, kind, mult, k, s, r in LAYERS:
        shp = shapes[f"net/{scope}/weights"]
        if kind == "conv":
            ho, wo = math.ceil(h / s), math.ceil(w / s)
            total += 2.0 * ho * wo * shp[3] * shp[2] * k * k
            h, w = ho, wo
        else:
            total += 2.0 * h * w * shp[3] * shp[2] * k * k
            h, w = h * 2, w * 2
    shp = shapes["net/color_pred/weights"]
    total += 2.0 * h * w * shp[2] * shp[3]
    return total / 1e9
