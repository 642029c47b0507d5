"""Drop-in ``MultiHeadSelfAttention`` (reference: deepfm/models/layers/attention.py:11-120).

Same constructor and module tree -- ``layers[i]`` with ``W_q, W_k, W_v: Linear(D, A)``,
``W_out: Linear(A, D)`` and ``layer_norm: LayerNorm(D)`` iff ``use_residual`` -- so ``state_dict``
keys and initialisation match; ``ValueError`` if ``attention_dim % num_heads`` (attention.py:41-44).
Each block's forward and backward are one fused shared-memory kernel (``dfm_attn_fwd`` /
``dfm_attn_bwd``); the backward recomputes the forward from the block input, so nothing but the
``(B, F, D)`` input is kept alive between the passes.
"""

from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import _lib


class _AttnBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, blk: "_AttentionBlock", x, *params):
        B, F, D = x.shape
        out = torch.empty_like(x)
        _lib.check(_lib.lib().dfm_attn_fwd(x.data_ptr(), B, F, D, blk.attention_dim, blk.num_heads,
                                           int(blk.use_residual), _lib.ptr_array(params), out.data_ptr(),
                                           _lib.stream_ptr()), "dfm_attn_fwd")
        ctx.blk = blk
        ctx.save_for_backward(x, *params)
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.lib()
        blk = ctx.blk
        x, *params = ctx.saved_tensors
        B, F, D = x.shape
        g_out = g_out.contiguous()
        g_x = torch.empty_like(x)
        g_params = [torch.empty_like(p) for p in params]
        nbytes = lib.dfm_attn_workspace_bytes(B, F, D, blk.attention_dim, blk.num_heads)
        ws = torch.empty((max(nbytes, 16),), device=x.device, dtype=torch.uint8)
        _lib.check(lib.dfm_attn_bwd(x.data_ptr(), g_out.data_ptr(), B, F, D, blk.attention_dim, blk.num_heads,
                                    int(blk.use_residual), _lib.ptr_array(params), g_x.data_ptr(),
                                    _lib.ptr_array(g_params), ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                   "dfm_attn_bwd")
        return (None, g_x, *g_params)


class _AttentionBlock(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int, attention_dim: int, use_residual: bool) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.attention_dim = attention_dim
        self.head_dim = attention_dim // num_heads
        self.scale = math.sqrt(self.head_dim)
        self.use_residual = use_residual
        self.W_q = nn.Linear(embed_dim, attention_dim)
        self.W_k = nn.Linear(embed_dim, attention_dim)
        self.W_v = nn.Linear(embed_dim, attention_dim)
        self.W_out = nn.Linear(attention_dim, embed_dim)
        if use_residual:
            self.layer_norm = nn.LayerNorm(embed_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        params = [self.W_q.weight, self.W_q.bias, self.W_k.weight, self.W_k.bias, self.W_v.weight,
                  self.W_v.bias, self.W_out.weight, self.W_out.bias]
        if self.use_residual:
            params += [self.layer_norm.weight, self.layer_norm.bias]
        for p in params:
            _lib.require_cuda(p, "attention parameter")
        return _AttnBlockFn.apply(self, x, *params)


class MultiHeadSelfAttention(nn.Module):
    def __init__(self, embed_dim: int, num_heads: int = 4, attention_dim: int = 64, num_layers: int = 1,
                 use_residual: bool = True) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.attention_dim = attention_dim
        self.head_dim = attention_dim // num_heads
        self.use_residual = use_residual
        if attention_dim % num_heads != 0:
            raise ValueError(f"attention_dim ({attention_dim}) must be divisible by num_heads ({num_heads})")
        self.layers = nn.ModuleList(
            _AttentionBlock(embed_dim, num_heads, attention_dim, use_residual) for _ in range(num_layers))

    def forward(self, field_embeddings: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(field_embeddings, "field_embeddings")
        if field_embeddings.dim() != 3 or field_embeddings.shape[2] != self.embed_dim:
            raise ValueError(f"attention expects (B, F, {self.embed_dim}), got {tuple(field_embeddings.shape)}")
        x = field_embeddings.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        for layer in self.layers:
            x = layer(x)
        return x
