"""Host-side mirror of the reference's ``matryodshka/nets.py`` for the MSI
inference path: ``msi_coord_train_net`` (nets.py:471-515).

The layer table below is the reference architecture; the arithmetic runs in the
sm_100a kernels of ``csrc/`` (tcgen05 implicit-GEMM convolutions, LayerNorm
reductions) through the C-ABI in ``include/msi_b200.h``.  There is no CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple


@dataclass(frozen=True)
class LayerSpec:
    scope: str          # TF variable scope (nets.py:486-507)
    kind: str           # "conv" (3x3 SAME + coord channel) | "deconv" (4x4 s2 SAME) | "head" (1x1 + bias + tanh)
    cout_mult: int      # Cout = ngf * cout_mult (head: num_outputs)
    k: int
    stride: int
    rate: int
    src: Tuple[str, ...]  # input feature scopes; two entries = channel concat (skip)


# nets.py:486-515.  "input" is the plane-sweep volume.
ARCH: List[LayerSpec] = [
    LayerSpec("conv1_1", "conv", 1, 3, 1, 1, ("input",)),
    LayerSpec("conv1_2", "conv", 2, 3, 2, 1, ("conv1_1",)),
    LayerSpec("conv2_1", "conv", 2, 3, 1, 1, ("conv1_2",)),
    LayerSpec("conv2_2", "conv", 4, 3, 2, 1, ("conv2_1",)),
    LayerSpec("conv3_1", "conv", 4, 3, 1, 1, ("conv2_2",)),
    LayerSpec("conv3_2", "conv", 4, 3, 1, 1, ("conv3_1",)),
    LayerSpec("conv3_3", "conv", 8, 3, 2, 1, ("conv3_2",)),
    LayerSpec("conv4_1", "conv", 8, 3, 1, 2, ("conv3_3",)),
    LayerSpec("conv4_2", "conv", 8, 3, 1, 2, ("conv4_1",)),
    LayerSpec("conv4_3", "conv", 8, 3, 1, 2, ("conv4_2",)),
    LayerSpec("conv6_1", "deconv", 4, 4, 2, 1, ("conv4_3", "conv3_3")),
    LayerSpec("conv6_2", "conv", 4, 3, 1, 1, ("conv6_1",)),
    LayerSpec("conv6_3", "conv", 4, 3, 1, 1, ("conv6_2",)),
    LayerSpec("conv7_1", "deconv", 2, 4, 2, 1, ("conv6_3", "conv2_2")),
    LayerSpec("conv7_2", "conv", 2, 3, 1, 1, ("conv7_1",)),
    LayerSpec("conv8_1", "deconv", 1, 4, 2, 1, ("conv7_2", "conv1_2")),
    LayerSpec("conv8_2", "conv", 1, 3, 1, 1, ("conv8_1",)),
    LayerSpec("color_pred", "head", 0, 1, 1, 1, ("conv8_2",)),
]


def layer_channels(num_inputs: int, num_outputs: int, ngf: int = 64) -> Dict[str, int]:
    ch = {"input": num_inputs}
    for l in ARCH:
        ch[l.scope] = num_outputs if l.kind == "head" else ngf * l.cout_mult
    return ch


def layer_shapes(num_inputs: int, num_outputs: int, ngf: int = 64, coord: bool = True) -> Dict[str, tuple]:
    """Weight shapes in the TF checkpoint layout, keyed by TF variable name
    (SURVEY.md 5): conv HWIO ``[k,k,Cin(+1 coord),Cout]``, deconv
    ``[k,k,Cout,Cin]``, LayerNorm gamma/beta ``[Cout]``, head bias."""
    ch = layer_channels(num_inputs, num_outputs, ngf)
    shapes: Dict[str, tuple] = {}
    for l in ARCH:
        cin = sum(ch[s] for s in l.src)
        cout = ch[l.scope]
        if l.kind == "conv":
            shapes[f"net/{l.scope}/weights"] = (l.k, l.k, cin + (1 if coord else 0), cout)
        elif l.kind == "deconv":
            shapes[f"net/{l.scope}/weights"] = (l.k, l.k, cout, cin)
        else:
            shapes[f"net/{l.scope}/weights"] = (1, 1, cin, cout)
            shapes[f"net/{l.scope}/biases"] = (cout,)
            continue
        shapes[f"net/{l.scope}/LayerNorm/gamma"] = (cout,)
        shapes[f"net/{l.scope}/LayerNorm/beta"] = (cout,)
    return shapes


def layer_geometry(H: int, W: int) -> Dict[str, Tuple[int, int]]:
    """Output (H, W) of every layer for an H x W input (SAME padding:
    ceil(n / stride); deconv doubles)."""
    hw = {"input": (H, W)}
    for l in ARCH:
        h, w = hw[l.src[0]]
        if l.kind == "deconv":
            hw[l.scope] = (h * 2, w * 2)
        else:
            hw[l.scope] = (-(-h // l.stride), -(-w // l.stride))
    return hw


def net_flops(H: int, W: int, num_inputs: int, num_outputs: int, ngf: int = 64, coord: bool = True) -> float:
    """FLOPs (2 x MACs) of one forward pass; SURVEY.md 8(a) a10 table formula
    (coord channels counted): 302.4 GFLOP at 320x640, P=L=32."""
    ch = layer_channels(num_inputs, num_outputs, ngf)
    hw = layer_geometry(H, W)
    total = 0.0
    for l in ARCH:
        cin = sum(ch[s] for s in l.src)
        cout = ch[l.scope]
        ho, wo = hw[l.scope]
        if l.kind == "conv":
            total += 2.0 * ho * wo * cout * (cin + (1 if coord else 0)) * l.k * l.k
        elif l.kind == "deconv":
            hi, wi = hw[l.src[0]]
            total += 2.0 * hi * wi * cin * cout * l.k * l.k
        else:
            total += 2.0 * ho * wo * cin * cout
    return total


def msi_coord_train_net(inputs, num_outputs, ngf=64, vscope="net", reuse_weights=False, *,
                        weights=None, engine=None):
    """nets.py:471 -- same positional signature as the reference.  ``inputs`` is
    the NHWC float32 plane-sweep volume on the GPU.  The reference finds its
    weights through TF variable scope ``vscope``; here they are an explicit dict
    keyed by the TF variable names (``weights=``), or a prepared ``engine``
    (``matryodshka_b200.runtime.NetEngine``) that already holds them packed."""
    from .runtime import NetEngine  # late import: runtime needs the CUDA library

    if engine is None:
        if weights is None:
            raise ValueError("msi_coord_train_net needs weights= (TF-named dict) or engine=")
        B, H, W, C = inputs.shape
        engine = NetEngine.cached(weights, H, W, C, num_outputs, ngf, inputs.device, vscope=vscope, max_batch=B)
    return _forward_op(engine, inputs)


def msi_train_net(inputs, num_outputs, ngf=64, vscope="net", reuse_weights=False, *, weights=None, engine=None):
    """nets.py:387-469 -- the non-coord net: every conv / deconv input goes through ``wrap_pad``
    (nets.py:288-295: circular padding along the width, zeros along the height) and a VALID conv;
    conv weights are ``[3,3,Cin,Cout]`` (no coord channel).  Same calling convention as
    ``msi_coord_train_net`` above."""
    from .runtime import NetEngine

    if engine is None:
        if weights is None:
            raise ValueError("msi_train_net needs weights= (TF-named dict) or engine=")
        B, H, W, C = inputs.shape
        engine = NetEngine.cached(weights, H, W, C, num_outputs, ngf, inputs.device, vscope=vscope, max_batch=B, variant="wrap")
    return _forward_op(engine, inputs)


def _forward_op(engine, inputs):
    """torch.ops.msi.net_forward on the engine's dispatcher handle."""
    import torch
    from . import torch_ops
    if getattr(engine, "_op_handle", None) is None:
        engine._op_handle = torch_ops.register_engine(engine)
    return torch.ops.msi.net_forward(inputs.contiguous().float(), engine._op_handle)
