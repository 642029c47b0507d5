"""Drop-in ``CIN`` (reference: deepfm/models/layers/cin.py:8-105).

Same constructor, attributes (``num_fields``, ``embed_dim``, ``split_half``, ``conv_layers`` --
``nn.Conv1d(K, L, kernel_size=1)`` kept as the parameter containers so ``state_dict`` keys and
initialisation are the reference's --, ``direct_sizes``, ``next_sizes``, ``output_dim``) and
``forward((B, F, D)) -> (B, output_dim)``.  Forward and backward run ``dfm_cin_fwd`` /
``dfm_cin_bwd``: the ``(B, H*F, D)`` outer product of cin.py:84-87 never exists in memory.
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _lib


class _CINFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: "CIN", x0, *wb):
        lib = _lib.lib()
        n = len(mod.conv_layers)
        weights, biases = wb[:n], wb[n:]
        B, F, D = x0.shape
        sizes = _lib.i32_array(mod.layer_sizes)
        info = (C.c_int64 * 4)()
        _lib.check(lib.dfm_cin_sizes(F, D, n, sizes, int(mod.split_half), B, info), "dfm_cin_sizes")
        out = torch.empty((B, info[0]), device=x0.device, dtype=torch.float32)
        acts = torch.empty((max(info[1] // 4, 1),), device=x0.device, dtype=torch.float32)
        prec = mod.precision_code_for(B)
        _lib.check(lib.dfm_cin_fwd(x0.data_ptr(), B, F, D, n, sizes, int(mod.split_half), _lib.ptr_array(weights),
                                   _lib.ptr_array(biases), prec, out.data_ptr(), acts.data_ptr(),
                                   _lib.stream_ptr()), "dfm_cin_fwd")
        ctx.mod, ctx.ws_bytes = mod, int(info[2])
        ctx.prec = prec
        ctx.save_for_backward(x0, acts, *weights)
        if mod.keep_activations:          # test aid: the post-ReLU activations (B, L_i, D) per layer, i.e. the ReLU decisions
            mod.last_activations, off = [], 0
            for L in mod.layer_sizes:
                mod.last_activations.append(acts[off: off + B * L * D].view(B, L, D))
                off += B * L * D
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.lib()
        mod = ctx.mod
        x0, acts, *weights = ctx.saved_tensors
        n = len(weights)
        B, F, D = x0.shape
        sizes = _lib.i32_array(mod.layer_sizes)
        g_out = g_out.contiguous()
        g_x0 = torch.empty_like(x0)
        g_w = [torch.empty_like(w) for w in weights]
        g_b = [torch.empty((w.shape[0],), device=w.device, dtype=torch.float32) for w in weights]
        ws = torch.empty((max(ctx.ws_bytes, 16),), device=x0.device, dtype=torch.uint8)
        _lib.check(lib.dfm_cin_bwd(x0.data_ptr(), g_out.data_ptr(), B, F, D, n, sizes, int(mod.split_half),
                                   _lib.ptr_array(weights), mod.backward_precision_code(ctx.prec), acts.data_ptr(), g_x0.data_ptr(),
                                   _lib.ptr_array(g_w), _lib.ptr_array(g_b), ws.data_ptr(), ws.numel(),
                                   _lib.stream_ptr()), "dfm_cin_bwd")
        return (None, g_x0, *g_w, *g_b)


class CIN(nn.Module):
    # "fp32": CUDA cores, reference-exact up to summation order.  "tf32": forward AND backward contractions on the
    # tcgen05 tensor cores (TF32 inputs, FP32 accumulate: 2e-3 forward, 3e-3 gradients on shared ReLU decisions).
    # "auto" (default): tf32 when the contraction is big enough to be tensor-bound (num_fields * widest layer >= 1024
    # and at least 4096 GEMM rows), fp32 -- the reference's arithmetic -- otherwise (ML-100K-sized models, tests).
    PRECISIONS = {"fp32": 0, "tf32": 1}

    def __init__(self, num_fields: int, embed_dim: int, layer_sizes: Optional[List[int]] = None,
                 split_half: bool = True) -> None:
        super().__init__()
        layer_sizes = list(layer_sizes or [128, 128])
        self.num_fields = num_fields
        self.embed_dim = embed_dim
        self.split_half = split_half
        self.layer_sizes = layer_sizes
        self.conv_layers = nn.ModuleList()
        self.direct_sizes: List[int] = []
        self.next_sizes: List[int] = []
        maps = num_fields
        last = len(layer_sizes) - 1
        for i, size in enumerate(layer_sizes):
            self.conv_layers.append(nn.Conv1d(maps * num_fields, size, kernel_size=1))
            if split_half and i < last:
                direct = size // 2
                self.direct_sizes.append(direct)
                self.next_sizes.append(size - direct)
                maps = size - direct
            else:
                self.direct_sizes.append(size)
                self.next_sizes.append(size)
                maps = size
        self.output_dim = sum(self.direct_sizes)
        self.precision = "auto"
        self.keep_activations = False
        self.last_activations = None

    def precision_code_for(self, batch: int) -> int:
        if self.precision == "auto":
            big = self.num_fields * max(self.layer_sizes) >= 1024 and batch * self.embed_dim >= 4096
            return 1 if big else 0
        return self.PRECISIONS[self.precision]

    def backward_precision_code(self, forward_code: int) -> int:
        """The precision is read again at backward time (tests back-propagate one graph at both precisions);
        "auto" follows what the forward chose."""
        return forward_code if self.precision == "auto" else self.PRECISIONS[self.precision]

    @property
    def precision_code(self) -> int:
        return self.PRECISIONS.get(self.precision, 0)

    def forward(self, field_embeddings: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(field_embeddings, "field_embeddings")
        if field_embeddings.dim() != 3 or field_embeddings.shape[1] != self.num_fields \
                or field_embeddings.shape[2] != self.embed_dim:
            raise ValueError(f"CIN expects (B, {self.num_fields}, {self.embed_dim}), got {tuple(field_embeddings.shape)}")
        x0 = field_embeddings.contiguous()
        if x0.dtype != torch.float32:
            x0 = x0.float()
        weights = [_lib.require_cuda(c.weight, "CIN weight") for c in self.conv_layers]
        biases = [c.bias for c in self.conv_layers]
        return _CINFn.apply(self, x0, *weights, *biases)
