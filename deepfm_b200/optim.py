"""Row-sparse Adam for the id tables (reference trainer.py:67-78 and :232-237).

The reference clips the global gradient norm to 1.0 and runs dense ``torch.optim.Adam`` over every parameter; at
Criteo scale that touches 7x the table bytes per step.  ``RowSparseAdam`` consumes the row-sparse gradients the
backward leaves in ``embedding.row_grads`` (sorted keys + per-segment sums, no compaction, no host sync) and
updates ``w``, ``exp_avg``, ``exp_avg_sq`` of the touched rows only (``dfm_adam_rows``); ``grad_sumsq()`` is the
tables' share of the global norm (``dfm_rows_sumsq``), so the caller can form the reference's clip coefficient
``min(1, max_norm / (norm + 1e-6))`` together with the dense parameters and pass it as ``clip_scale``.
Semantics: torch.optim.Adam restricted to touched rows ("lazy" moments; oracle: ``adam_rows``).  Works for
``FeatureEmbedding`` and for ``ShardedFeatureEmbedding`` (each rank steps the rows it owns; replicated small tables
and every other parameter stay with the caller's dense optimizer).
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib


class RowSparseAdam:
    def __init__(self, embedding, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
        self.emb, self.lr, self.betas, self.eps = embedding, float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        self._params = embedding._ordered_params()
        self._is_table = list(embedding._param_is_table)
        for p, t in zip(self._params, self._is_table):
            if t:
                _lib.require_cuda(p, "embedding table")
        self.exp_avg = [torch.zeros_like(p) if t else None for p, t in zip(self._params, self._is_table)]
        self.exp_avg_sq = [torch.zeros_like(p) if t else None for p, t in zip(self._params, self._is_table)]
        self._ssq_ws = None

    def _plan(self):
        if hasattr(self.emb, "_ensure_plans"):          # sharded: the owner-side plan
            return self.emb._ensure_plans()[0]
        return self.emb._ensure_plan()

    def _slots(self, tensors) -> C.Array:
        n_fields = len(self.emb.field_names)
        arr = (C.c_void_p * (5 * n_fields))()
        slots = getattr(self.emb, "_slot_of_param", None)
        if slots is None or len(slots) != len(tensors):
            self.emb._ordered_params()
            slots = self.emb._slot_of_param
        for slot, t, tab in zip(slots, tensors, self._is_table):
            arr[slot] = t.data_ptr() if (tab and t is not None) else None
        return arr

    def table_parameters(self) -> List[torch.Tensor]:
        return [p for p, t in zip(self._params, self._is_table) if t]

    def grad_sumsq(self) -> torch.Tensor:
        """Device scalar: sum of squares of the touched rows' gradients (both views) of the last backward."""
        rg = self.emb.row_grads
        out = torch.zeros((1,), device=self._params[0].device, dtype=torch.float32)
        if rg is None:
            return out
        lib = _lib.lib()
        ws = torch.empty((lib.dfm_rows_sumsq_workspace_bytes(),), device=out.device, dtype=torch.uint8)
        _lib.check(lib.dfm_rows_sumsq(self._plan(), rg.sorted_keys.numel(), rg.sorted_keys.data_ptr(), rg.row_grad2.data_ptr(),
                                      rg.row_grad1.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                   "dfm_rows_sumsq")
        return out

    @torch.no_grad()
    def step(self, clip_scale: Optional[torch.Tensor] = None) -> None:
        rg = self.emb.row_grads
        if rg is None:
            raise RuntimeError("RowSparseAdam.step: no row-sparse gradients (embedding.grad_mode must be 'row_sparse' "
                               "and backward() must have run)")
        self.step_count += 1
        lib = _lib.lib()
        # incrementally maintained ||W||^2 (layers/l2.py): if the cached value describes the tables as they are now,
        # the kernel adds sum(w_new^2 - w_old^2) of the rows it touches and the cache follows the new versions;
        # otherwise the next get_l2_reg_loss() falls back to the exact reduction.
        tables = self.table_parameters()
        cache = getattr(self.emb, "_l2_cache", None)
        live = cache is not None and cache.valid_for(tables)
        acc = ws = None
        if live:
            acc = cache.acc
            if self._ssq_ws is None or self._ssq_ws.device != acc.device:
                self._ssq_ws = torch.empty((lib.dfm_adam_rows_workspace_bytes(),), device=acc.device, dtype=torch.uint8)
            ws = self._ssq_ws
        _lib.check(lib.dfm_adam_rows(self._plan(), rg.sorted_keys.numel(), rg.sorted_keys.data_ptr(), rg.row_grad2.data_ptr(),
                                     rg.row_grad1.data_ptr(), self._slots(self._params), self._slots(self.exp_avg),
                                     self._slots(self.exp_avg_sq), self.lr, self.betas[0], self.betas[1], self.eps,
                                     self.step_count, _lib.ptr(clip_scale), _lib.ptr(acc), _lib.ptr(ws),
                                     ws.numel() if ws is not None else 0, _lib.stream_ptr()), "dfm_adam_rows")
        for p, t in zip(self._params, self._is_table):      # the kernel wrote through raw pointers
            if t:
                torch.autograd.graph.increment_version(p)
        if live:
            cache.rekey(tables)
