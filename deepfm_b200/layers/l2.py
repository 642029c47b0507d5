"""The embedding L2 penalty ``lambda * sum_p ||p||^2`` (reference: base.py:78-83) as one autograd node.

Value.  The reference re-reads every parameter every step (a Python loop of ``norm().pow()``); at the Criteo shape that
is an 8.6 GB pass for a scalar that only changes on the rows the optimizer touched (SURVEY hard part 2).  Here the
id tables' share ``sum ||w||^2`` is kept in a device fp64 scalar (``TableNormCache``):
  * computed once by the exact fixed-order reduction ``dfm_sumsq_acc`` (the audited slow path),
  * reused for as long as the table tensors are the same storage at the same ``_version`` (nobody wrote to them),
  * updated in place by ``RowSparseAdam.step`` -- ``dfm_adam_rows`` adds ``sum(w_new^2 - w_old^2)`` of the rows it
    touches -- so a training loop never pays the dense pass again;
  * any other writer (a dense torch optimizer, ``load_state_dict`` ...) bumps ``_version`` and the next call falls
    back to the exact reduction.
The small parameters (DENSE-field Linears, projections) are reduced every call (a few KB).

Gradient.  If the penalty is added to a loss that also flows through the ``FeatureEmbedding`` forward of the same
step, the ``2*lambda*p`` term is folded into the embedding backward kernel K2 (no extra pass over the tables).  The
node takes a scalar output of the embedding's autograd node as an (otherwise unused) input, so the autograd graph itself
guarantees that the embedding's backward runs after this node in every backward pass that reaches it -- no engine
callbacks, no private API.  (The anchor is a dedicated scalar output of the embedding node that the module holds on
to: the three views themselves may be dropped by the model as soon as it has combined them.)  In every other situation the node produces ``2*lambda*p`` itself with ``dfm_axpy``.
"""

from __future__ import annotations

from typing import List, Optional

import torch

from .. import _lib


class TableNormCache:
    """Device fp64 ``sum ||w||^2`` over a fixed list of table tensors, valid for one (storage, version) key."""

    def __init__(self) -> None:
        self.key = None
        self.acc: Optional[torch.Tensor] = None      # (1,) float64 on the tables' device
        self.refreshes = 0                           # number of exact (dense) reductions so far

    @staticmethod
    def key_of(tables) -> tuple:
        return tuple(sorted((p.data_ptr(), p._version, p.numel()) for p in tables))      # a set: callers list the tables in different orders

    def valid_for(self, tables) -> bool:
        return self.acc is not None and self.key == self.key_of(tables)

    def refresh(self, tables) -> torch.Tensor:
        """Exact reduction (one pass over the tables) on the current stream."""
        lib = _lib.lib()
        dev = tables[0].device
        if self.acc is None or self.acc.device != dev:
            self.acc = torch.zeros((1,), device=dev, dtype=torch.float64)
        ws = torch.empty((4096,), device=dev, dtype=torch.float32)
        _lib.check(lib.dfm_sumsq_acc(len(tables), _lib.ptr_array(tables), _lib.i64_array([p.numel() for p in tables]),
                                     self.acc.data_ptr(), ws.data_ptr(), _lib.stream_ptr()), "dfm_sumsq_acc")
        self.key = self.key_of(tables)
        self.refreshes += 1
        return self.acc

    def rekey(self, tables) -> None:
        """The caller updated ``acc`` itself while writing to the tables (RowSparseAdam): adopt the new versions."""
        self.key = self.key_of(tables)

    def invalidate(self) -> None:
        self.key = None


def _split_params(emb):
    """(tables, small): the id tables (cached) and everything else (reduced every call).  The split itself is cached on
    the module (cleared by its ``_apply``, i.e. by ``.to()`` / ``.cuda()``): two module-tree walks per step otherwise."""
    cached = getattr(emb, "_l2_split", None)
    if cached is not None:
        return cached[0], cached[1]
    params = [p for p in emb.parameters()]
    table_ids = set()
    ordered = getattr(emb, "_ordered_params", None)
    if ordered is not None:      # exactly the tensors RowSparseAdam steps (sharded module: the sharded tables only)
        op = ordered()
        table_ids = {id(p) for p, is_table in zip(op, emb._param_is_table) if is_table}
    tables = [p for p in params if id(p) in table_ids]
    small = [p for p in params if id(p) not in table_ids]
    if hasattr(emb, "_ordered_cache"):       # only modules that invalidate the cache in _apply
        emb._l2_split = (tables, small, params)
    return tables, small


def table_norm_cache(emb) -> TableNormCache:
    cache = getattr(emb, "_l2_cache", None)
    if cache is None:
        cache = emb._l2_cache = TableNormCache()
    return cache


def _l2_value(emb, params, lam: float) -> torch.Tensor:
    lib = _lib.lib()
    dev = params[0].device
    tables, small = _split_params(emb)
    out = torch.empty((), device=dev, dtype=torch.float32)
    acc_t = None
    if tables:
        cache = table_norm_cache(emb)
        acc_t = cache.acc if cache.valid_for(tables) else cache.refresh(tables)
    acc_s = None
    if small:
        acc_s = torch.empty((1,), device=dev, dtype=torch.float64)
        ws = torch.empty((4096,), device=dev, dtype=torch.float32)
        _lib.check(lib.dfm_sumsq_acc(len(small), _lib.ptr_array(small), _lib.i64_array([p.numel() for p in small]),
                                     acc_s.data_ptr(), ws.data_ptr(), _lib.stream_ptr()), "dfm_sumsq_acc")
    a, b = (acc_t, acc_s) if acc_t is not None else (acc_s, None)
    _lib.check(lib.dfm_l2_combine(a.data_ptr(), _lib.ptr(b), float(lam), out.data_ptr(), _lib.stream_ptr()), "dfm_l2_combine")
    return out


def prefetch_l2(emb, lam: float) -> None:
    """Kept for API compatibility: the table norm is cached / maintained incrementally now, nothing to prefetch."""
    return None


class _L2PenaltyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, lam: float, anchor, *params):
        out = _l2_value(emb, params, lam)
        ctx.emb, ctx.lam = emb, float(lam)
        # `anchor` is an output of the embedding's autograd node of this step (or None): with it as an input, that
        # node is an ancestor of this one and is guaranteed to run later in the same backward pass.
        live = emb._live_ctx() if getattr(emb, "_live_ctx", None) is not None else None
        ctx.live = live if anchor is not None else None
        ctx.params = params           # inputs of this node (leaf parameters): plain references, not ~90 saved tensors
        return out

    @staticmethod
    def backward(ctx, g):
        params = ctx.params
        emb, lam = ctx.emb, ctx.lam
        g = g.contiguous().float()
        live = ctx.live
        if live is not None and live.l2 is None:
            live.l2 = (lam, g)       # K2 adds 2*lam*g*p to every gradient it writes; consumed (cleared) there
            return (None, None, None) + (None,) * len(params)
        grads = []
        lib = _lib.lib()
        for p in params:
            gp = torch.empty_like(p)
            _lib.check(lib.dfm_axpy(p.data_ptr(), p.numel(), 2.0 * lam, g.data_ptr(), gp.data_ptr(), 0,
                                    _lib.stream_ptr()), "dfm_axpy")
            grads.append(gp)
        return (None, None, None) + tuple(grads)


def l2_penalty(emb, lam: float) -> torch.Tensor:
    cached = getattr(emb, "_l2_split", None)
    if cached is None:
        _split_params(emb)
        cached = getattr(emb, "_l2_split", None)
    params: List[torch.Tensor] = [p.contiguous() for p in (cached[2] if cached is not None else emb.parameters())]
    for p in params:
        _lib.require_cuda(p, "FeatureEmbedding parameter")
    anchor = getattr(emb, "_live_anchor", None)
    live = emb._live_ctx() if getattr(emb, "_live_ctx", None) is not None else None
    if anchor is not None and (live is None or not torch.is_grad_enabled() or not anchor.requires_grad):
        anchor = None
    return _L2PenaltyFn.apply(emb, lam, anchor, *params)
