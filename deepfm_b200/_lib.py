"""ctypes binding of ``libdeepfm_b200.so`` (the C ABI declared in ``include/deepfm_b200.h``).

There is no CPU / eager fallback: if the library is missing or a call fails, the caller gets a
``RuntimeError`` naming the problem.  ``python -m deepfm_b200.build`` (or
``__graft_entry__.build()``) compiles the library in-tree with nvcc for sm_100a.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdeepfm_b200.so")

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4
SPARSE, SEQUENCE, DENSE = 0, 1, 2
SUM, MEAN, MAX = 0, 1, 2
GRAD_DENSE, GRAD_ROWSPARSE, GRAD_SKIP_TABLES = 0, 1, 2
GRAD_PRESORTED = 0x100
KIND = {"sparse": SPARSE, "sequence": SEQUENCE, "dense": DENSE}
COMBINER = {"sum": SUM, "mean": MEAN, "max": MAX}

_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_pp = C.POINTER(C.c_void_p)
_pi32, _pi64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)

# name -> (restype, argtypes); kept in one table so tests can check every header symbol is exported
SIGNATURES = {
    "dfm_last_error": (C.c_char_p, []),
    "dfm_version": (C.c_int, []),
    "dfm_device_info": (C.c_int, [C.c_int, _pi64]),
    "dfm_plan_create": (_vp, [C.c_int, _pi32, _pi32, _pi64, _pi32, _pi32, C.c_int]),
    "dfm_plan_destroy": (None, [_vp]),
    "dfm_plan_info": (C.c_int, [_vp, _pi64]),
    "dfm_plan_slots": (C.c_int, [_vp, _pi32, _pi32, _pi64]),
    "dfm_embed_fwd": (C.c_int, [_vp, _i64, _pp, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "dfm_embed_bwd_workspace_bytes": (_sz, [_vp, _i64]),
    "dfm_embed_bwd": (C.c_int, [_vp, _i64, _pp, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _f32, _vp, C.c_int, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dfm_emit_keys": (C.c_int, [_vp, _i64, _pp, _vp, _vp]),
    "dfm_sort_keys_workspace_bytes": (_sz, [_vp, _i64]),
    "dfm_sort_keys": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dfm_fm_fwd": (C.c_int, [_vp, _i64, C.c_int, C.c_int, _vp, _vp]),
    "dfm_fm_bwd": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, _vp, _vp]),
    "dfm_sumsq": (C.c_int, [C.c_int, _pp, _pi64, _f32, _vp, _vp, _vp]),
    "dfm_sumsq_acc": (C.c_int, [C.c_int, _pp, _pi64, _vp, _vp, _vp]),
    "dfm_l2_combine": (C.c_int, [_vp, _vp, _f32, _vp, _vp]),
    "dfm_axpy": (C.c_int, [_vp, _i64, _f32, _vp, _vp, C.c_int, _vp]),
    "dfm_cin_sizes": (C.c_int, [C.c_int, C.c_int, C.c_int, _pi32, C.c_int, _i64, _pi64]),
    "dfm_cin_fwd": (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, _pi32, C.c_int, _pp, _pp, C.c_int,
                              _vp, _vp, _vp]),
    "dfm_cin_bwd": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, C.c_int, _pi32, C.c_int, _pp, C.c_int,
                              _vp, _vp, _pp, _pp, _vp, _sz, _vp]),
    "dfm_shard_route_workspace_bytes": (_sz, [_vp, _i64]),
    "dfm_shard_route": (C.c_int, [_vp, C.c_int, _pi64, _i64, _pp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dfm_shard_gather": (C.c_int, [_vp, C.c_int, C.c_int, _pi64, _i64, _vp, _pp, _vp, _vp, _vp]),
    "dfm_shard_pack_grad": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp]),
    "dfm_shard_gather_p2p": (C.c_int, [_vp, C.c_int, C.c_int, _pi64, _i64, _vp, _pp, C.c_int, _pi64, _pp, _vp, _vp]),
    "dfm_shard_pack_grad_p2p": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, _pi64, _pp, _vp, _f32, _vp]),
    "dfm_plan_set_table_stride": (C.c_int, [_vp, C.c_int, C.c_int]),
    "dfm_plan_set_field_source": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dfm_rows_bwd_workspace_bytes": (_sz, [_vp, _i64]),
    "dfm_rows_bwd": (C.c_int, [_vp, _i64, _pp, _vp, _vp, _vp, _f32, _vp, C.c_int, _pp, _vp, _vp, _vp, _vp, _vp,
                               _vp, _sz, _vp]),
    "dfm_shard_ukeys": (C.c_int, [_vp, C.c_int, _pi64, _pi64, C.c_int, _i64, _pp, _vp, _vp, _vp, _vp, _vp]),
    "dfm_sort_pairs_workspace_bytes": (_sz, [_i64, C.c_int]),
    "dfm_sort_pairs": (C.c_int, [_i64, C.c_int, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dfm_shard_unique_workspace_bytes": (_sz, [_i64]),
    "dfm_shard_unique": (C.c_int, [_vp, C.c_int, C.c_int, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dfm_shard_gather2": (C.c_int, [_vp, C.c_int, C.c_int, _i64, _vp, _pp, C.c_int, _pi64, _pp, _pp, _vp, _vp]),
    "dfm_shard_push_counts": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
    "dfm_shard_push_keys": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _i64, _vp, _vp]),
    "dfm_tower_store_floats": (C.c_size_t, [C.c_int, _vp, _i64, _vp]),
    "dfm_tower_seq_workspace_bytes": (C.c_size_t, [C.c_int, _vp, _i64, C.c_int]),
    "dfm_tower_fwd": (C.c_int, [C.c_int, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_float, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dfm_tower_bwd": (C.c_int, [C.c_int, _vp, _i64, _vp, _vp, _vp, _vp, C.c_int, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "dfm_peer_barrier": (C.c_int, [_vp, C.c_int, C.c_int, C.c_uint32, _vp]),
    "dfm_shard_bwd_peer": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, C.c_uint32, C.c_int,
                                     _pi64, _pp, _pp, _f32, _vp, _sz, _vp]),
    "dfm_adam_rows": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _pp, _pp, _pp, _f32, _f32, _f32, _f32, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dfm_adam_rows_workspace_bytes": (_sz, []),
    "dfm_rows_sumsq_workspace_bytes": (_sz, []),
    "dfm_rows_sumsq": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dfm_gemm3_workspace_bytes": (_sz, [C.c_int, _i64, _i64, _i64]),
    "dfm_gemm3": (C.c_int, [C.c_int, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "dfm_tower_workspace_bytes": (_sz, [_i64, C.c_int]),
    "dfm_bn_stats": (C.c_int, [_vp, _i64, C.c_int, _f32, _vp, _vp, _vp, _vp, _f32, _vp, _sz, _vp]),
    "dfm_bn_act_fwd": (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _f32, C.c_uint64, _vp, _vp]),
    "dfm_bn_act_bwd": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _f32, C.c_uint64,
                                 _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dfm_head_fwd": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, _vp, _vp]),
    "dfm_head_bwd": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dfm_attn_workspace_bytes": (_sz, [_i64, C.c_int, C.c_int, C.c_int, C.c_int]),
    "dfm_attn_fwd": (C.c_int, [_vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _pp, _vp, _vp]),
    "dfm_attn_bwd": (C.c_int, [_vp, _vp, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _pp, _vp, _pp,
                               _vp, _sz, _vp]),
}

_lock = threading.Lock()
_lib = None


def header_symbols() -> list[str]:
    """Function names declared in include/deepfm_b200.h (parsed, so the test cannot drift)."""
    import re
    text = open(os.path.join(HERE, "..", "include", "deepfm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dfm_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    """Load the library once; raise loudly if it is not there (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the deepfm_b200 modules have no CPU or eager fallback. "
                "Build it with `python -m deepfm_b200.build` (needs nvcc, targets sm_100a).")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                continue            # symbols are validated against the header by the tests
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().dfm_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc == OK:
        return
    msg = last_error()
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(f"{what}: {msg}")
    if rc == ERR_INVALID:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None stays NULL)."""
    return None if t is None else t.data_ptr()


def ptr_array(tensors) -> C.Array:
    arr = (C.c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else (t if isinstance(t, int) else t.data_ptr())
    return arr


def i32_array(values) -> C.Array:
    return (C.c_int32 * max(len(values), 1))(*[int(v) for v in values])


def i64_array(values) -> C.Array:
    return (C.c_int64 * max(len(values), 1))(*[int(v) for v in values])


_raw_stream = None


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device.  Called ~40 times per step: the raw-handle getter
    (what torch's own extensions use) costs ~1 us, ``torch.cuda.current_stream().cuda_stream`` ~15 us; the public API is
    the fallback if the getter is ever renamed."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        fn = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        _raw_stream = fn if fn is not None else False
    if _raw_stream:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name: str):
    """The hot path runs on a B200 only: refuse anything that is not a CUDA tensor."""
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: deepfm_b200 kernels are CUDA (sm_100a) only and have no "
            "CPU/MPS fallback; move the module and the batch to a CUDA device")
    return t
