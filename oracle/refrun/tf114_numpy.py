"""NumPy stand-in for the slice of TensorFlow 1.14 (+ tf.contrib.slim, tensorflow_graphics 1.0) that the
reference's inference path calls -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Purpose: execute the REFERENCE'S OWN SOURCE FILES (/root/reference/geometry/*.py,
matryodshka/{msi,nets}.py, unmodified, imported from where they lie) in this container, where
TensorFlow 1.14 cannot be installed, and record what they compute (oracle/refrun/run_reference.py ->
tests/golden/reference_run.npz).  The restatement in oracle/*.py and the CUDA path are then pinned to
those vectors instead of to a reading of the source.

What is restated here is only the third-party op layer (tensorflow==1.14.0, tensorflow-graphics==1.0.0,
matryodshka-gpu.yml:14,43-44), eagerly, in float32 with one rounding per op:

* tensors are ``numpy.ndarray`` views with ``get_shape().as_list()``; Python scalars meeting a tensor
  are rounded once to the tensor's dtype (NumPy >= 2 weak-scalar promotion == TF's constant conversion);
* ``tf.linspace``: the 1.14 CPU kernel, ``start + step * i`` with ``step = (stop - start) / (num - 1)`` in T;
* ``tf.matmul``: sum over k in index order, separately rounded products and sums;
* ``tf.mod`` floor-mod; ``tf.add_n`` left to right; ``tf.gather_nd``; ``tf.image.convert_image_dtype``
  (uint8 -> float: ``x * (1/255)``; float -> uint8: truncating ``cast(x * 255.5)``, no saturation);
  ``tf.image.resize`` (bilinear, align_corners=True, the 1.14 legacy scaler);
* ``slim.conv2d`` / ``conv2d_transpose`` (SAME / VALID, stride, rate; bias only without a normalizer),
  ``slim.layer_norm`` (``nn.moments`` over axes 1..3, ``nn.batch_normalization`` with eps 1e-12),
  ``slim.arg_scope``; variables are read from ``set_variables({name: array})`` by their checkpoint names;
  convolutions are an im2col + float32 matmul written here (independent of torch);
* ``tfgt.rotation_matrix_3d.from_euler``: R = Rz Ry Rx from the sines / cosines of the angles.

Nothing here is imported by the product or by the restatement in oracle/*.py.
"""
from __future__ import annotations

import contextlib
import functools
import sys
import types
import warnings

import numpy as np

warnings.filterwarnings("ignore", category=RuntimeWarning)  # sqrt of negatives, x / inf: same IEEE results as TF

float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
uint8 = np.uint8
bool = np.bool_  # noqa: A001  (tf.bool)


class TensorShape(list):
    def as_list(self):
        return list(self)


class Tensor(np.ndarray):
    """ndarray view with the two TF-1.x accessors the reference uses."""

    def get_shape(self):
        return TensorShape(self.shape)


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, str):
        return np.dtype(dtype).type
    return dtype


def _t(x, dtype=None):
    return np.asarray(x, dtype=_dt(dtype)).view(Tensor)


def _default_float(value, dtype):
    """tf.constant / convert_to_tensor: Python floats (and float64 NumPy data without a dtype) -> float32
    only when the value is a Python object; NumPy arrays keep their dtype."""
    if dtype is not None:
        return _t(value, dtype)
    if isinstance(value, np.ndarray):
        return _t(value)
    a = np.asarray(value)
    if a.dtype == np.float64:
        return _t(a, np.float32)
    if a.dtype == np.int64:
        return _t(a, np.int32)
    return _t(a)


# ---- construction -----------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    t = _default_float(value, dtype)
    if shape is not None:
        shape = [int(v) for v in shape]
        t = t.reshape(shape) if t.size == int(np.prod(shape)) else np.broadcast_to(t, shape)
    return _t(t)


def convert_to_tensor(value, dtype=None, name=None):
    return _default_float(value, dtype)


def zeros(shape, dtype=float32, name=None):
    return _t(np.zeros(shape, _dt(dtype)))


def ones(shape, dtype=float32, name=None):
    return _t(np.ones(shape, _dt(dtype)))


def zeros_like(x, dtype=None):
    return _t(np.zeros_like(np.asarray(x), dtype=_dt(dtype)))


def ones_like(x, dtype=None):
    return _t(np.ones_like(np.asarray(x), dtype=_dt(dtype)))


def eye(n, dtype=float32, name=None):
    return _t(np.eye(n, dtype=_dt(dtype)))


def range(*args, **kw):  # noqa: A001
    return _t(np.arange(*args), np.int32)


def linspace(start, stop, num, name=None):
    """[TF-1.14 LinSpaceOp, core/kernels/sequence_ops.cc]: flat(i) = start + step * i, all in T = float32."""
    start = np.float32(start)
    stop = np.float32(stop)
    if num == 1:
        return _t(np.array([start], np.float32))
    step = np.float32((stop - start) / np.float32(num - 1))
    i = np.arange(num).astype(np.float32)
    return _t((start + step * i).astype(np.float32))


def meshgrid(*args, **kw):
    return [_t(a) for a in np.meshgrid(*[np.asarray(a) for a in args], indexing=kw.get("indexing", "xy"))]


def is_tensor(x):
    return isinstance(x, np.ndarray)


def identity(x, name=None):
    return _t(x)


# ---- shape ops -----------------------------------------------------------------------------------------
def shape(x, name=None):
    return list(np.asarray(x).shape)


def reshape(x, shape, name=None):  # noqa: A002
    return _t(np.reshape(np.asarray(x), [int(s) for s in shape]))


def expand_dims(x, axis=None, name=None, dim=None):
    return _t(np.expand_dims(np.asarray(x), axis if axis is not None else dim))


def tile(x, multiples, name=None):
    return _t(np.tile(np.asarray(x), multiples))


def transpose(x, perm=None, name=None):
    return _t(np.transpose(np.asarray(x), perm))


def matrix_transpose(x, name=None):
    return _t(np.swapaxes(np.asarray(x), -1, -2))


def concat(values, axis, name=None):
    return _t(np.concatenate([np.asarray(v) for v in values], axis=axis))


def stack(values, axis=0, name=None):
    return _t(np.stack([np.asarray(v) for v in values], axis=axis))


def unstack(x, num=None, axis=0, name=None):
    x = np.asarray(x)
    return [_t(np.take(x, i, axis=axis)) for i in np.arange(x.shape[axis])]


def slice(x, begin, size, name=None):  # noqa: A001
    import builtins
    idx = tuple(builtins.slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return _t(np.asarray(x)[idx])


def pad(x, paddings, mode="CONSTANT", name=None, constant_values=0):
    assert mode == "CONSTANT"
    return _t(np.pad(np.asarray(x), paddings, mode="constant", constant_values=constant_values))


def gather(params, indices, axis=0, name=None):
    return _t(np.take(np.asarray(params), np.asarray(indices), axis=axis))


def gather_nd(params, indices, name=None):
    params = np.asarray(params)
    indices = np.asarray(indices)
    k = indices.shape[-1]
    return _t(params[tuple(indices[..., i] for i in np.arange(k))])


def cast(x, dtype, name=None):
    """tf.cast float -> int32 truncates toward zero (C cast), like ndarray.astype."""
    return _t(np.asarray(x).astype(_dt(dtype)))


# ---- elementwise (float32 in, float32 out, one rounding) ---------------------------------------------
def _unary(fn):
    def op(x, name=None):
        return _t(fn(np.asarray(x)))
    return op


cos = _unary(np.cos)
sin = _unary(np.sin)
sqrt = _unary(np.sqrt)
floor = _unary(np.floor)
abs = _unary(np.abs)  # noqa: A001
sign = _unary(np.sign)
is_nan = _unary(np.isnan)
tanh = _unary(np.tanh)


def square(x, name=None):
    x = np.asarray(x)
    return _t(x * x)


def atan2(y, x, name=None):
    return _t(np.arctan2(np.asarray(y), np.asarray(x)))


def mod(x, y, name=None):
    """tf.mod == floormod."""
    return _t(np.mod(np.asarray(x), y))


def divide(x, y, name=None):
    return _t(np.asarray(x) / y)


def _cmp(fn):
    def op(x, y, name=None):
        return _t(fn(np.asarray(x), y))
    return op


greater = _cmp(np.greater)
greater_equal = _cmp(np.greater_equal)
less = _cmp(np.less)
less_equal = _cmp(np.less_equal)
equal = _cmp(np.equal)


def where(condition, x=None, y=None, name=None):
    return _t(np.where(np.asarray(condition), np.asarray(x), np.asarray(y)))


def add_n(inputs, name=None):
    acc = np.asarray(inputs[0])
    for v in inputs[1:]:
        acc = acc + np.asarray(v)
    return _t(acc)


def reduce_mean(x, axis=None, keepdims=False, name=None, keep_dims=None):
    if keep_dims is not None:
        keepdims = keep_dims
    return _t(np.mean(np.asarray(x), axis=tuple(axis) if isinstance(axis, (list, tuple)) else axis, keepdims=keepdims,
                      dtype=np.asarray(x).dtype))


def matmul(a, b, name=None):
    """[..., M, K] x [..., K, N]: sum over k in index order; every product and sum rounded to float32."""
    a = np.asarray(a)
    b = np.asarray(b)
    acc = a[..., :, 0:1] * b[..., 0:1, :]
    for k in np.arange(1, a.shape[-1]):
        acc = acc + a[..., :, k:k + 1] * b[..., k:k + 1, :]
    return _t(acc)


def matrix_inverse(x, name=None):
    return _t(np.linalg.inv(np.asarray(x, np.float64)).astype(np.float32))


# ---- scopes, graph, flags ------------------------------------------------------------------------------
_SCOPE = []


@contextlib.contextmanager
def name_scope(name, *a, **kw):
    yield


@contextlib.contextmanager
def variable_scope(name, reuse=None, **kw):
    _SCOPE.append(name)
    try:
        yield
    finally:
        _SCOPE.pop()


class _Graph:
    def __init__(self):
        self.tensors = {}

    def get_tensor_by_name(self, name):
        return self.tensors[name]


_GRAPH = _Graph()


def get_default_graph():
    return _GRAPH


def set_named_tensor(name, value):
    """What test.py does with tf.matrix_inverse(..., name='ref_pose_inv') (test.py:111)."""
    _GRAPH.tensors[name] = _t(value, np.float32)


class _Flags:
    def __getattr__(self, k):
        raise AttributeError("FLAGS.%s was not set by the runner" % k)


def _noop(*a, **kw):
    return None


flags_mod = types.SimpleNamespace(FLAGS=_Flags(), DEFINE_string=_noop, DEFINE_integer=_noop, DEFINE_float=_noop,
                                  DEFINE_boolean=_noop, DEFINE_bool=_noop)
app = types.SimpleNamespace(flags=flags_mod)
FLAGS = flags_mod.FLAGS


def set_flags(**kw):
    for k, v in kw.items():
        object.__setattr__(flags_mod.FLAGS, k, v)


# ---- tf.image ------------------------------------------------------------------------------------------
def _convert_image_dtype(image, dtype, saturate=False, name=None):
    """[TF-1.14 image_ops_impl.convert_image_dtype]."""
    image = np.asarray(image)
    dtype = _dt(dtype)
    if image.dtype == dtype:
        return _t(image)
    if image.dtype == np.uint8 and dtype == np.float32:
        return _t(image.astype(np.float32) * np.float32(1.0 / 255))
    if image.dtype == np.float32 and dtype == np.uint8:
        assert not saturate
        scaled = image * np.float32(255 + 0.5)
        return _t(scaled.astype(np.int32).astype(np.uint8))
    raise NotImplementedError((image.dtype, dtype))


def _resize_bilinear(images, size, method=0, align_corners=False, name=None):
    """[TF-1.14 tf.image.resize / ResizeBilinear kernel, align_corners=True, legacy scaler]: scale = (in - 1) /
    (out - 1) in float32; in = out_index * scale; lower = floor(in), upper = min(ceil(in), in - 1), lerp = in - lower;
    top = tl + (tr - tl) * x_lerp; bottom = bl + (br - bl) * x_lerp; out = top + (bottom - top) * y_lerp."""
    assert method == 0 and align_corners, "only BILINEAR with align_corners=True is on the path (test.py:319-325)"
    x = np.asarray(images, np.float32)
    B, h, w, C = x.shape
    oh, ow = int(size[0]), int(size[1])

    def axis(n_in, n_out):
        scale = np.float32((n_in - 1) / np.float32(n_out - 1)) if n_out > 1 else np.float32(0)
        pos = np.arange(n_out).astype(np.float32) * scale
        lo = np.floor(pos)
        hi = np.minimum(np.ceil(pos), n_in - 1)
        return lo.astype(np.int64), hi.astype(np.int64), (pos - lo).astype(np.float32)

    y0, y1, ly = axis(h, oh)
    x0, x1, lx = axis(w, ow)
    out = np.empty((B, oh, ow, C), np.float32)
    for i in np.arange(oh):      # written row by row, independently of oracle/highres_np.py
        tl, tr = x[:, y0[i]][:, x0], x[:, y0[i]][:, x1]
        bl, br = x[:, y1[i]][:, x0], x[:, y1[i]][:, x1]
        top = tl + (tr - tl) * lx[None, :, None]
        bottom = bl + (br - bl) * lx[None, :, None]
        out[:, i] = top + (bottom - top) * ly[i]
    return _t(out)


image = types.SimpleNamespace(convert_image_dtype=_convert_image_dtype, resize=_resize_bilinear,
                              resize_images=_resize_bilinear,
                              ResizeMethod=types.SimpleNamespace(BILINEAR=0, NEAREST_NEIGHBOR=1))


# ---- tf.nn + tf.contrib.slim ---------------------------------------------------------------------------
def _relu(x, name=None):
    x = np.asarray(x)
    return _t(np.maximum(x, np.float32(0)))


def _same_pads(n, k_eff, s):
    """[TF SAME]: out = ceil(n / s); pad_total = max((out - 1) * s + k_eff - n, 0); before = pad_total // 2."""
    out = -(-n // s)
    total = max((out - 1) * s + k_eff - n, 0)
    return total // 2, total - total // 2


def _conv2d_raw(x, w, stride, rate, padding):
    """x [B,H,W,Ci] float32, w [kh,kw,Ci,Co]; im2col + one float32 matmul."""
    x = np.asarray(x, np.float32)
    w = np.asarray(w, np.float32)
    kh, kw, ci, co = w.shape
    assert x.shape[3] == ci, (x.shape, w.shape)
    if padding == "SAME":
        pt, pb = _same_pads(x.shape[1], (kh - 1) * rate + 1, stride)
        pl, pr = _same_pads(x.shape[2], (kw - 1) * rate + 1, stride)
        x = np.pad(x, [(0, 0), (pt, pb), (pl, pr), (0, 0)])
    else:
        assert padding == "VALID"
    B, H, W, _ = x.shape
    oh = (H - ((kh - 1) * rate + 1)) // stride + 1
    ow = (W - ((kw - 1) * rate + 1)) // stride + 1
    cols = np.empty((B, oh, ow, kh, kw, ci), np.float32)
    for ky in np.arange(kh):
        for kx in np.arange(kw):
            y0, x0 = ky * rate, kx * rate
            cols[:, :, :, ky, kx, :] = x[:, y0:y0 + (oh - 1) * stride + 1:stride, x0:x0 + (ow - 1) * stride + 1:stride, :]
    out = cols.reshape(B * oh * ow, kh * kw * ci) @ w.reshape(kh * kw * ci, co)
    return out.reshape(B, oh, ow, co).astype(np.float32)


def _conv2d_transpose_raw(x, w, stride, padding):
    """[TF conv2d_transpose = gradient of conv2d w.r.t. its input], w [kh,kw,Co,Ci]:
    full[b, i*s + ky, j*s + kx, o] += x[b,i,j,c] * w[ky,kx,o,c]; VALID keeps the full (n-1)s + k extent,
    SAME crops it to n*s starting at the forward conv's pad_before."""
    x = np.asarray(x, np.float32)
    w = np.asarray(w, np.float32)
    kh, kw, co, ci = w.shape
    B, H, W, _ = x.shape
    assert x.shape[3] == ci, (x.shape, w.shape)
    fh, fw = (H - 1) * stride + kh, (W - 1) * stride + kw
    full = np.zeros((B, fh, fw, co), np.float32)
    flat = x.reshape(B * H * W, ci)
    for ky in np.arange(kh):
        for kx in np.arange(kw):
            contrib = (flat @ w[ky, kx].T).reshape(B, H, W, co)
            full[:, ky:ky + (H - 1) * stride + 1:stride, kx:kx + (W - 1) * stride + 1:stride, :] += contrib
    if padding == "VALID":
        return full
    oh, ow = H * stride, W * stride
    pt, _ = _same_pads(oh, kh, stride)
    pl, _ = _same_pads(ow, kw, stride)
    return np.ascontiguousarray(full[:, pt:pt + oh, pl:pl + ow, :])


_VARIABLES = {}
variables_read = []


def set_variables(d):
    _VARIABLES.clear()
    _VARIABLES.update({k: np.asarray(v, np.float32) for k, v in d.items()})
    del variables_read[:]


def _get_variable(name):
    full = "/".join(_SCOPE + [name])
    variables_read.append(full)
    return _VARIABLES[full]


_ARGS = [{}]


@contextlib.contextmanager
def _arg_scope(fns, **kw):
    new = dict(_ARGS[-1])
    for f in fns:
        d = dict(new.get(f.__name__, {}))
        d.update(kw)
        new[f.__name__] = d
    _ARGS.append(new)
    try:
        yield
    finally:
        _ARGS.pop()


def _scoped(fn):
    @functools.wraps(fn)
    def wrapper(*a, **kw):
        merged = dict(_ARGS[-1].get(fn.__name__, {}))
        merged.update(kw)
        return fn(*a, **merged)
    return wrapper


def _pair(v):
    if isinstance(v, (list, tuple)):
        assert v[0] == v[1]
        return int(v[0])
    return int(v)


@_scoped
def layer_norm(inputs, center=True, scale=True, activation_fn=None, scope=None, begin_norm_axis=1, begin_params_axis=-1,
               **kw):
    """[slim.layer_norm, TF 1.14]: nn.moments over axes 1.. (two-pass variance), nn.batch_normalization with
    variance_epsilon = 1e-12: inv = rsqrt(var + eps) * gamma; y = x * inv + (beta - mean * inv)."""
    with variable_scope(scope or "LayerNorm"):
        beta = _get_variable("beta")
        gamma = _get_variable("gamma")
    x = np.asarray(inputs, np.float32)
    axes = tuple(np.arange(1, x.ndim))
    mean = np.mean(x, axis=axes, keepdims=True, dtype=np.float32)
    d = x - mean
    var = np.mean(d * d, axis=axes, keepdims=True, dtype=np.float32)
    inv = (np.float32(1) / np.sqrt(var + np.float32(1e-12))) * gamma
    y = x * inv + (beta - mean * inv)
    if activation_fn is not None:
        y = activation_fn(y)
    return _t(y.astype(np.float32))


def _finish(out, normalizer_fn, normalizer_params, activation_fn, co):
    if normalizer_fn is not None:
        out = normalizer_fn(out, **(normalizer_params or {}))
    else:
        out = np.asarray(out) + _get_variable("biases").reshape(1, 1, 1, co)
    if activation_fn is not None:
        out = activation_fn(out)
    return _t(np.asarray(out, np.float32))


@_scoped
def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", rate=1, activation_fn=_relu, normalizer_fn=None,
           normalizer_params=None, scope=None, **kw):
    """[slim.conv2d, TF 1.14]: weights [kh,kw,Cin,Cout]; biases only when normalizer_fn is None."""
    assert scope is not None
    with variable_scope(scope):
        w = _get_variable("weights")
        k = _pair(kernel_size)
        assert w.shape[0] == k and w.shape[1] == k and w.shape[3] == num_outputs, (scope, w.shape, k, num_outputs)
        out = _conv2d_raw(inputs, w, _pair(stride), _pair(rate), padding)
        return _finish(out, normalizer_fn, normalizer_params, activation_fn, num_outputs)


@_scoped
def conv2d_transpose(inputs, num_outputs, kernel_size, stride=1, padding="SAME", activation_fn=_relu, normalizer_fn=None,
                     normalizer_params=None, scope=None, **kw):
    """[slim.conv2d_transpose, TF 1.14]: weights [kh,kw,Cout,Cin]."""
    assert scope is not None
    with variable_scope(scope):
        w = _get_variable("weights")
        k = _pair(kernel_size)
        assert w.shape[0] == k and w.shape[1] == k and w.shape[2] == num_outputs, (scope, w.shape, k, num_outputs)
        out = _conv2d_transpose_raw(inputs, w, _pair(stride), padding)
        return _finish(out, normalizer_fn, normalizer_params, activation_fn, num_outputs)


nn = types.SimpleNamespace(relu=_relu, tanh=tanh)


# ---- tensorflow_graphics.geometry.transformation.rotation_matrix_3d ------------------------------------
def _from_euler(angles, name=None):
    """[tensorflow_graphics 1.0 rotation_matrix_3d.from_euler]: angles [..., 3] = (x, y, z), R = Rz Ry Rx."""
    a = np.asarray(_default_float(angles, None), np.float32)
    s, c = np.sin(a), np.cos(a)
    sx, sy, sz = s[..., 0], s[..., 1], s[..., 2]
    cx, cy, cz = c[..., 0], c[..., 1], c[..., 2]
    m = np.stack([cy * cz, (sx * sy * cz) - (cx * sz), (cx * sy * cz) + (sx * sz),
                  cy * sz, (sx * sy * sz) + (cx * cz), (cx * sy * sz) - (sx * cz),
                  -sy, sx * cy, cx * cy], axis=-1)
    return _t(m.reshape(a.shape[:-1] + (3, 3)).astype(np.float32))


# ---- installation --------------------------------------------------------------------------------------
def install():
    """Registers this module as `tensorflow` (+ contrib.slim, tensorflow_graphics) in sys.modules and stubs the
    two imports of matryodshka/msi.py that are off the inference path (E-LPIPS loss, PNG writer)."""
    me = sys.modules[__name__]
    slim = types.ModuleType("tensorflow.contrib.slim")
    slim.conv2d = conv2d
    slim.conv2d_transpose = conv2d_transpose
    slim.layer_norm = layer_norm
    slim.arg_scope = _arg_scope
    contrib = types.ModuleType("tensorflow.contrib")
    contrib.slim = slim
    me.contrib = contrib
    me.Tensor = Tensor
    sys.modules["tensorflow"] = me
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.slim"] = slim
    tfg = types.ModuleType("tensorflow_graphics")
    tfg_geo = types.ModuleType("tensorflow_graphics.geometry")
    tfg_tr = types.ModuleType("tensorflow_graphics.geometry.transformation")
    tfg_tr.rotation_matrix_3d = types.SimpleNamespace(from_euler=_from_euler)
    tfg.geometry = tfg_geo
    tfg_geo.transformation = tfg_tr
    sys.modules["tensorflow_graphics"] = tfg
    sys.modules["tensorflow_graphics.geometry"] = tfg_geo
    sys.modules["tensorflow_graphics.geometry.transformation"] = tfg_tr
    elp = types.ModuleType("elpips")
    elp_in = types.ModuleType("elpips.elpips")
    elp.elpips = elp_in
    sys.modules["elpips"] = elp
    sys.modules["elpips.elpips"] = elp_in
    utils = types.ModuleType("matryodshka.utils")
    utils.write_image = _noop
    sys.modules["matryodshka.utils"] = utils
    return me
