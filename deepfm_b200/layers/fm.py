"""Drop-in ``FMInteraction`` (reference: deepfm/models/layers/fm.py:9-23).

Parameter-free; ``forward((B,F,D)) -> (B,1)`` = ``0.5 * sum_d((sum_f e)^2 - sum_f e^2)``.
When the input is the ``field_embeddings`` tensor a ``FeatureEmbedding`` forward just returned,
the value was already produced by the fused kernel K1 (the embeddings are read once); otherwise
the stand-alone kernels ``dfm_fm_fwd`` / ``dfm_fm_bwd`` run.
"""

from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib


class _FMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e):
        B, F, D = e.shape
        out = torch.empty((B, 1), device=e.device, dtype=torch.float32)
        _lib.check(_lib.lib().dfm_fm_fwd(e.data_ptr(), B, F, D, out.data_ptr(), _lib.stream_ptr()), "dfm_fm_fwd")
        ctx.save_for_backward(e)
        return out

    @staticmethod
    def backward(ctx, g):
        (e,) = ctx.saved_tensors
        B, F, D = e.shape
        ge = torch.empty_like(e)
        _lib.check(_lib.lib().dfm_fm_bwd(e.data_ptr(), g.contiguous().data_ptr(), B, F, D, ge.data_ptr(),
                                         _lib.stream_ptr()), "dfm_fm_bwd")
        return ge


class FMInteraction(nn.Module):
    def forward(self, field_embeddings: torch.Tensor) -> torch.Tensor:
        cached = getattr(field_embeddings, "_dfm_fm", None)
        if cached is not None and cached[1] == field_embeddings._version:
            return cached[0]
        _lib.require_cuda(field_embeddings, "field_embeddings")
        if field_embeddings.dim() != 3:
            raise ValueError(f"FMInteraction expects (B, F, D), got {tuple(field_embeddings.shape)}")
        e = field_embeddings.contiguous()
        if e.dtype != torch.float32:
            e = e.float()
        return _FMFn.apply(e)
