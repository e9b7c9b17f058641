"""Builds libmsi_b200.so (the C-ABI library of include/msi_b200.h) in-tree with nvcc for sm_100a.

    python -m matryodshka_b200.build [--force]

nvcc cross-compiles without a GPU.  The geometry translation unit is compiled with
-fmad=false (strict IEEE float32, see csrc/geom_device.cuh); the rest with default FMA
contraction.  The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmsi_b200.so")
OBJ = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math",
          "--expt-relaxed-constexpr"]
# translation unit -> extra flags
UNITS = {
    "api_common.cu": [],
    "geom_kernels.cu": ["-fmad=false"],
    "layernorm.cu": [],
    "conv_simt.cu": [],
    "conv_tcgen05.cu": [],
    "net.cu": [],
}


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "msi_b200.h"))
    return hdrs


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force: bool = False, verbose: bool = True) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _deps()
    objs = []
    procs = []
    for unit, extra in UNITS.items():
        src = os.path.join(CSRC, unit)
        obj = os.path.join(OBJ, unit.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for unit, p in procs:
        out, _ = p.communicate()
        if out.strip() and verbose:
            print(out)
        if p.returncode != 0:
            print(out, file=sys.stderr)
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(OUT, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", OUT, *objs, "-lcudart", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(OUT)
