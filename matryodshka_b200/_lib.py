"""ctypes binding of the C ABI in include/msi_b200.h (libmsi_b200.so, built in-tree by
``python -m matryodshka_b200.build``).

There is no CPU fallback: a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSI_B200_LIB") or os.path.join(_HERE, "libmsi_b200.so")  # env: A/B builds of the library

MSI_OK = 0
IMG_F32, IMG_U8 = 0, 1
CONV_TCGEN05, CONV_SIMT = 0, 1
PREC_FP16X3, PREC_FP16, PREC_FP16_FP8X = 0, 1, 2
NET_COORD, NET_WRAP = 0, 1
COLOR_MODES = {"blend_psv": 0, "blend_bg": 1, "blend_bg_psv": 2, "alpha_only": 3}
ACT_SCALE = 16.0
ABI_VERSION = 1

# name -> (restype, argtypes); every symbol include/msi_b200.h declares
_P = c_void_p
_I = c_int
SIGNATURES = {
    "msi_b200_abi_version": (c_int, []),
    "msi_last_error": (c_char_p, []),
    "msi_launch_count": (c_uint64, []),
    "msi_psv_build": (c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P,
                              c_size_t, _P]),
    "msi_psv_scratch_bytes": (c_size_t, [_I, _I, _I]),
    "msi_sweep_table_bytes": (c_size_t, [_I, _I, _I, _I]),
    "msi_sweep_table_build": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "msi_psv_gather": (c_int, [_P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, c_size_t, _P]),
    "msi_sweep_coords": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "msi_rgba_assemble": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "msi_rgba_assemble_ex": (c_int, [_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "msi_rgba_assemble_strided": (c_int, [_P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "msi_render_composite": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "msi_render_composite_gather": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P,
                                            ctypes.c_longlong, _P]),
    "msi_render_ods": (c_int, [_P, _P, ctypes.c_float, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "msi_render_perspective": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "msi_intersect_sphere_coords": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "msi_intersect_sphere_coords_ex": (c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "msi_project_layers": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "msi_point_op": (c_int, [_I, _P, _P, _P, ctypes.c_longlong, _I, _P, _I, ctypes.c_float, ctypes.c_float, _I, _I,
                             _P, _P, _P, _P, _P]),
    "msi_resample": (c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "msi_over_composite": (c_int, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "msi_highres_plane": (c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P]),
    "msi_highres_composite": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "msi_net_create": (c_int, [POINTER(c_void_p), _I, _I, _I, _I, _I, _I, _I, _I]),
    "msi_net_create_ex": (c_int, [POINTER(c_void_p), _I, _I, _I, _I, _I, _I, _I, _I, _I]),
    "msi_net_destroy": (None, [_P]),
    "msi_net_workspace_bytes": (c_size_t, [_P]),
    "msi_net_arena_bytes": (c_size_t, [_P]),
    "msi_net_input_c_stride": (c_int, [_P]),
    "msi_net_bind": (c_int, [_P, _P, c_size_t, _P, c_size_t]),
    "msi_net_load_layer": (c_int, [_P, c_char_p, _P, _P, _P, _P, _P]),
    "msi_net_forward": (c_int, [_P, _P, _P, _P, _I, _P, _P]),
    "msi_net_can_fuse_rgba": (c_int, [_P]),
    "msi_net_forward_rgba": (c_int, [_P, _P, _P, _P, _I, _P, _P]),
    "msi_net_input_buffers": (c_int, [_P, POINTER(c_void_p), POINTER(c_void_p)]),
    "msi_net_read_activation": (c_int, [_P, c_char_p, _I, _P, _P]),
    "msi_net_read_raw": (c_int, [_P, c_char_p, _I, _P, _P]),
    "msi_net_num_launches_per_forward": (c_int, [_P]),
    "msi_net_forward_profiled": (c_int, [_P, _P, _P, _P, _I, _P, _P, POINTER(ctypes.c_float), POINTER(ctypes.c_float)]),
    "msi_net_forward_profiled_flush": (c_int, [_P, _P, _P, _P, _I, _P, _P, POINTER(ctypes.c_float), POINTER(ctypes.c_float),
                                               _P, c_size_t, _P]),
    "msi_net_num_layers": (c_int, [_P]),
    "msi_net_layer_scope": (c_char_p, [_P, _I]),
    "msi_net_layer_flops": (ctypes.c_double, [_P, _I]),
}

_lib = None


class MsiError(RuntimeError):
    pass


def load():
    """dlopen the library (no GPU needed for this step) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MsiError(
            f"{LIB_PATH} is missing: build the CUDA library first (python -m matryodshka_b200.build). "
            "matryodshka_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    if lib.msi_b200_abi_version() != ABI_VERSION:
        raise MsiError("libmsi_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise MsiError("matryodshka_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def check(rc: int, what: str = ""):
    if rc != MSI_OK:
        msg = load().msi_last_error()
        raise MsiError(f"{what or 'msi_b200 call'} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise MsiError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise MsiError("expected a contiguous tensor")
    return c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    return int(load().msi_launch_count())
