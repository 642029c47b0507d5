"""Optional: run the (out-of-scope, library) DNN GEMMs on cuBLAS 12.9's FP32 emulation.

The MLP tower of the models stays ``nn.Linear`` -> cuBLAS sgemm (not a custom kernel).  PyTorch
2.11+cu128 bundles cuBLAS 12.8, whose fp32 GEMM on B200 is a CUDA-core (SIMT) kernel at ~55 TFLOP/s; the CUDA
12.9 toolkit in this image ships cuBLAS 12.9, which can compute the same fp32 GEMM on the BF16 tensor cores as
nine BF16 products of three-way splits of the operands with fp32 accumulation ("BF16x9",
``CUBLAS_EMULATE_SINGLE_PRECISION``).  That mode is at least as accurate as the native fp32 kernel (measured on
the bench shapes: max-norm relative error vs fp64 4.4e-7, native 1.8e-6) and 1.2-2.3x faster.

``enable()`` must run BEFORE ``import torch``: it loads the toolkit's libcublasLt / libcublas with RTLD_GLOBAL,
so that torch binds to them instead of its bundled 12.8 copy (same SONAME, ABI-compatible minor update), and
sets the emulation environment variables.  It is a no-op (returns False) when torch is already imported or the
toolkit libraries are absent.  Nothing in the hot-path kernels of this package depends on it.
"""

from __future__ import annotations

import ctypes
import os
import sys

TOOLKIT_LIBS = ("/usr/local/cuda/lib64/libcublasLt.so.12", "/usr/local/cuda/lib64/libcublas.so.12")
_state = {"enabled": False, "why": "not requested"}


def enable(strategy: str = "performant") -> bool:
    if _state["enabled"]:
        return True
    if "torch" in sys.modules:
        _state["why"] = "torch was imported first (its bundled cuBLAS is already bound)"
        return False
    if not all(os.path.exists(p) for p in TOOLKIT_LIBS):
        _state["why"] = "CUDA 12.9 toolkit cuBLAS not found"
        return False
    try:
        for p in TOOLKIT_LIBS:
            ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
    except OSError as e:
        _state["why"] = f"could not load the toolkit cuBLAS: {e}"
        return False
    os.environ["CUBLAS_EMULATE_SINGLE_PRECISION"] = "1"
    os.environ.setdefault("CUBLAS_EMULATION_STRATEGY", strategy)
    _state["enabled"], _state["why"] = True, "cuBLAS 12.9 FP32 emulation (BF16x9)"
    return True


def status() -> str:
    return _state["why"]


def cublas_version() -> str:
    """Version of the libcublas the process is bound to (after torch is imported)."""
    try:
        lib = ctypes.CDLL("libcublas.so.12")
        v, out = ctypes.c_int(), []
        for i in range(3):
            lib.cublasGetProperty(i, ctypes.byref(v))
            out.append(str(v.value))
        return ".".join(out)
    except Exception:
        return "unknown"
