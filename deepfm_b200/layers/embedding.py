"""Drop-in ``FeatureEmbedding`` running on the fused sm_100a kernels K1 / K2.

Mirrors the reference module (deepfm/models/layers/embedding.py:14-126): same constructor,
attributes (``schema``, ``fm_embed_dim``, ``field_names``, ``second_order_embeddings``,
``first_order_embeddings``, ``projections``), ``state_dict`` keys, parameter enumeration order and
initialisation, and the same ``forward(batch) -> (first_order, field_embeddings, flat)``.
The per-field ``nn.Embedding`` / ``nn.EmbeddingBag`` / ``nn.Linear`` children are kept as
parameter containers only -- their ATen forwards are never called; forward and backward are one
``torch.autograd.Function`` over ``dfm_embed_fwd`` / ``dfm_embed_bwd``.

Extras that do not exist in the reference (all default to reference semantics):
  * ``grad_mode``: ``"dense"`` (reference: dense ``(V, d)`` table gradients, ``2*l2*w`` on every
    row) or ``"row_sparse"`` (Criteo scale: table gradients stay as sorted unique rows in
    ``self.row_grads``; L2 applied to touched rows only; see ``RowSparseGrads``).
  * the FM term is computed by the same kernel and handed to ``FMInteraction`` through the
    returned ``field_embeddings`` tensor, so DeepFM reads the embeddings from HBM once.
"""

from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib
from ..schema import kind_of
from ._status import IndexStatusMixin


class RowSparseGrads:
    """Table gradients of one backward pass in sorted-segment form (no host sync to build).

    ``sorted_keys[p]`` is the global row (``row_base[field] + id``) at sorted position ``p``;
    positions ``>= n_valid`` hold the PAD key.  At the first position of every run of equal keys
    ``row_grad2[p, :d_field]`` / ``row_grad1[p]`` hold that row's summed gradient (+ ``2*l2*w``).
    ``counts`` is a device int64[2] = (n_valid, n_unique).
    """

    def __init__(self, sorted_keys, sorted_payload, row_grad2, row_grad1, counts, row_base, dims, names):
        self.sorted_keys, self.sorted_payload = sorted_keys, sorted_payload
        self.row_grad2, self.row_grad1, self.counts = row_grad2, row_grad1, counts
        self.row_base, self.dims, self.names = row_base, dims, names

    def heads(self) -> torch.Tensor:
        """Sorted positions that start a segment (one host sync)."""
        k = self.sorted_keys
        n_valid = int(self.counts[0].item())
        k = k[:n_valid]
        if n_valid == 0:
            return torch.empty(0, dtype=torch.long, device=k.device)
        head = torch.ones(n_valid, dtype=torch.bool, device=k.device)
        head[1:] = k[1:] != k[:-1]
        return head.nonzero(as_tuple=True)[0]

    def per_table(self) -> Dict[str, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        """{field: (local_rows int64 (U,), grad2 (U, d), grad1 (U,))} -- the unique touched rows."""
        pos = self.heads()
        keys = self.sorted_keys[pos].long()
        out = {}
        for f, name in enumerate(self.names):
            lo, hi = self.row_base[f], self.row_base[f + 1]
            if hi == lo:
                continue
            m = (keys >= lo) & (keys < hi)
            p = pos[m]
            out[name] = (keys[m] - lo, self.row_grad2[p, : self.dims[f]], self.row_grad1[p])
        return out


class _EmbedFn(torch.autograd.Function):
    """forward = K1 (dfm_embed_fwd), backward = K2 (dfm_embed_bwd)."""

    @staticmethod
    def forward(ctx, mod: "FeatureEmbedding", n_inputs: int, need_bwd: bool, *tensors):
        inputs, params = tensors[:n_inputs], tensors[n_inputs:]
        lib = _lib.lib()
        dev = params[0].device
        B = inputs[0].shape[0]
        F, D, T, S, A = mod.num_fields, mod.fm_embed_dim, mod._T, mod._S, mod._A
        flat = torch.empty((B, T), device=dev, dtype=torch.float32)
        field = flat.view(B, F, D) if mod._aliasable else torch.empty((B, F, D), device=dev, dtype=torch.float32)
        first = torch.empty((B, 1), device=dev, dtype=torch.float32)
        fm_out = torch.empty((B, 1), device=dev, dtype=torch.float32)
        fm_sum = torch.empty((B, D), device=dev, dtype=torch.float32) if need_bwd else None
        prep = mod._take_prepared(inputs) if (need_bwd and S) else None     # keys sorted ahead of the step (prepare())
        keys = torch.empty((B * S,), device=dev, dtype=torch.int32) if (need_bwd and S and prep is None) else None
        aux = torch.empty((B, A), device=dev, dtype=torch.int32) if A else None
        status = mod._status_word(dev)
        in_arr = _lib.ptr_array(inputs)
        par_arr = mod._param_ptrs(params)
        ev = mod._event_pair("fwd")
        _lib.check(lib.dfm_embed_fwd(mod._plan, B, in_arr, par_arr, first.data_ptr(), field.data_ptr(),
                                     flat.data_ptr(), fm_out.data_ptr(), _lib.ptr(fm_sum), _lib.ptr(keys),
                                     _lib.ptr(aux), _lib.ptr(status), _lib.stream_ptr()), "dfm_embed_fwd")
        if ev is not None:
            ev[1].record()
        if status is not None:
            mod._post_status()
        if keys is not None and mod.async_sort and B * S > 0:
            # the backward's sort needs the keys only: start it now on a side stream, underneath the interaction /
            # DNN forward and backward; the embedding backward then begins at the segmented reduction
            prep = mod._sort_async(keys, B * S, dev)
        ctx.mod = mod
        ctx.n_inputs = n_inputs
        ctx.set_materialize_grads(False)
        ctx.l2 = None            # (lambda, upstream-grad tensor) set by the L2 penalty node, consumed by backward
        ctx.prep = prep
        if need_bwd:
            # outputs / intermediates are saved tensors; the batch tensors and the parameters are INPUTS of this node and
            # stay plain references (packing and unpacking ~130 saved tensors costs ~0.2 ms of host time per step)
            ctx.save_for_backward(field, flat, fm_sum, keys, aux)
            ctx.inputs, ctx.params = inputs, params
            mod._live_ctx = weakref.ref(ctx)
        # 5th output: a scalar that exists only to be an input of the L2 penalty node (layers/l2.py), which makes this
        # node an ancestor of that one -- the module keeps it alive, the three views may be dropped by the caller
        anchor = torch.zeros((), device=dev, dtype=torch.float32)
        return first, field, flat, fm_out, anchor

    @staticmethod
    def backward(ctx, g_first, g_field, g_flat, g_fm, _g_anchor=None):
        mod: FeatureEmbedding = ctx.mod
        mod._live_anchor = None           # a penalty node created from now on cannot hang below this (spent) node
        mod.raise_if_bad_index(block=False)   # lazy check of the forward's status word (the reference raises IndexError)
        lib = _lib.lib()
        field, flat, fm_sum, keys, aux = ctx.saved_tensors
        inputs, params = ctx.inputs, ctx.params
        dev = flat.device
        B = flat.shape[0]
        S = mod._S
        N = B * S
        cont = lambda g: None if g is None else g.contiguous()
        g_first, g_field, g_flat, g_fm = cont(g_first), cont(g_field), cont(g_flat), cont(g_fm)
        rowsparse = mod.grad_mode == "row_sparse"
        grads: List[Optional[torch.Tensor]] = []
        for p, is_table in zip(params, mod._param_is_table):
            grads.append(None if (rowsparse and is_table) else torch.empty_like(p))
        lam, gscale = (0.0, None)
        if ctx.l2 is not None:
            lam, gscale = ctx.l2
            ctx.l2 = None         # consumed: a second backward over a retained graph gets its own fold (or none)
        ws_bytes = lib.dfm_embed_bwd_workspace_bytes(mod._plan, B)
        ws = torch.empty((max(ws_bytes, 16),), device=dev, dtype=torch.uint8)
        prep = ctx.prep
        flags = 0
        if prep is not None:          # sorted by the input pipeline on its own stream: nothing to sort here
            skeys, spay, ev_sorted = prep
            torch.cuda.current_stream(dev).wait_event(ev_sorted)
            skeys.record_stream(torch.cuda.current_stream(dev))
            spay.record_stream(torch.cuda.current_stream(dev))
            flags = _lib.GRAD_PRESORTED
        else:
            skeys = torch.empty((max(N, 1),), device=dev, dtype=torch.int32)
            spay = torch.empty((max(N, 1),), device=dev, dtype=torch.int32)
        counts = torch.zeros((2,), device=dev, dtype=torch.int64)
        rg2 = rg1 = None
        if rowsparse:
            rg2 = torch.empty((max(N, 1), max(mod._max_tdim, 1)), device=dev, dtype=torch.float32)
            rg1 = torch.empty((max(N, 1),), device=dev, dtype=torch.float32)
        ev = mod._event_pair("bwd")
        _lib.check(lib.dfm_embed_bwd(
            mod._plan, B, _lib.ptr_array(inputs), mod._param_ptrs(params),
            _lib.ptr(g_first), _lib.ptr(g_field), _lib.ptr(g_flat), _lib.ptr(g_fm),
            field.data_ptr(), flat.data_ptr(), _lib.ptr(fm_sum), _lib.ptr(keys), _lib.ptr(aux),
            float(lam), _lib.ptr(gscale), (_lib.GRAD_ROWSPARSE if rowsparse else _lib.GRAD_DENSE) | flags,
            mod._param_ptrs(grads), skeys.data_ptr(), spay.data_ptr(), _lib.ptr(rg2), _lib.ptr(rg1),
            counts.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "dfm_embed_bwd")
        if ev is not None:
            ev[1].record()
        if rowsparse:
            mod.row_grads = RowSparseGrads(skeys, spay, rg2, rg1, counts, mod._row_base, mod._dims, mod.field_names)
        else:
            mod.row_grads = None
        mod.last_counts = counts
        return (None, None, None) + (None,) * ctx.n_inputs + tuple(grads)


class FeatureEmbedding(IndexStatusMixin, nn.Module):
    """Three views from one fused pass: ``first_order (B,1)``, ``field_embeddings (B,F,D)``,
    ``flat_embeddings (B,T)``  (reference: embedding.py:14-126)."""

    def __init__(self, schema, fm_embed_dim: int = 16) -> None:
        super().__init__()
        self.schema = schema
        self.fm_embed_dim = fm_embed_dim
        self.field_names = list(schema.fields.keys())
        self.second_order_embeddings = nn.ModuleDict()
        self.first_order_embeddings = nn.ModuleDict()
        self.projections = nn.ModuleDict()
        kinds, dims, vocabs, lens, combs = [], [], [], [], []
        for name in self.field_names:
            fs = schema.fields[name]
            kind = kind_of(fs)
            if kind not in _lib.KIND:
                raise ValueError(f"field {name!r}: unknown feature type {kind!r}")
            d = int(fs.embedding_dim)
            if kind == "sparse":
                second = nn.Embedding(fs.vocabulary_size, d, padding_idx=0)
                first = nn.Embedding(fs.vocabulary_size, 1, padding_idx=0)
            elif kind == "sequence":
                if fs.combiner not in _lib.COMBINER:
                    raise ValueError(f"field {name!r}: unknown combiner {fs.combiner!r}")
                second = nn.EmbeddingBag(fs.vocabulary_size, d, mode=fs.combiner, padding_idx=0)
                first = nn.EmbeddingBag(fs.vocabulary_size, 1, mode=fs.combiner, padding_idx=0)
            else:
                second, first = nn.Linear(1, d), nn.Linear(1, 1)
            self.second_order_embeddings[name] = second
            self.first_order_embeddings[name] = first
            if d != fm_embed_dim:
                self.projections[name] = nn.Linear(d, fm_embed_dim, bias=False)
            kinds.append(_lib.KIND[kind])
            dims.append(d)
            vocabs.append(int(fs.vocabulary_size) if kind != "dense" else 0)
            lens.append(int(fs.max_length) if kind == "sequence" else 1)
            combs.append(_lib.COMBINER[fs.combiner] if kind == "sequence" else _lib.SUM)
        self._kinds, self._dims = kinds, dims
        self._init_weights()

        # extras
        self.grad_mode = "dense"
        self.profile_events = None      # set to {} to collect (start, end) CUDA events per call
        self._init_status()             # out-of-range ids raise IndexError lazily (layers/_status.py)
        self.async_sort = True          # sort the backward's keys on a side stream right behind K1 (False: inside backward)
        self._live_anchor = None
        self.row_grads: Optional[RowSparseGrads] = None
        self.last_counts = None
        self._live_ctx = None
        self._plan = None
        self._plan_args = (kinds, dims, vocabs, lens, combs)
        self.num_fields = len(self.field_names)
        # static derived sizes (pure host arithmetic, identical to dfm_plan_info)
        self._T = sum(dims)
        self._S = sum(l for k, l in zip(kinds, lens) if k != _lib.DENSE)
        self._A = sum((1 if c == _lib.MEAN else d + 1 if c == _lib.MAX else 0)
                      for k, d, c in zip(kinds, dims, combs) if k == _lib.SEQUENCE)
        self._aliasable = all(d == fm_embed_dim for d in dims)
        self._max_tdim = max([d for k, d in zip(kinds, dims) if k != _lib.DENSE], default=0)
        rb = [0]
        for k, v in zip(kinds, vocabs):
            rb.append(rb[-1] + (v if k != _lib.DENSE else 0))
        self._row_base = rb
        # parameter order handed to the kernels: per field (w2, b2, w1, b1, proj)
        self._param_is_table: List[bool] = []
        self._slot_of_param: List[int] = []

    # -- reference: embedding.py:66-74 ------------------------------------------------------
    def _init_weights(self) -> None:
        for m in self.modules():
            if isinstance(m, (nn.Embedding, nn.EmbeddingBag)):
                nn.init.xavier_uniform_(m.weight.data[1:])      # row 0 (padding) stays zero
            elif isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight.data)
                if m.bias is not None:
                    nn.init.zeros_(m.bias.data)

    # -- measurement hook: CUDA events around the C-ABI calls on the launching stream -----------
    def _event_pair(self, which: str):
        store = getattr(self, "profile_events", None)
        if store is None:
            return None
        pair = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        store.setdefault(which, []).append(pair)
        pair[0].record()
        return pair

    # -- kernel plumbing -----------------------------------------------------------------------
    def _ensure_plan(self):
        if self._plan is None:
            lib = _lib.lib()
            kinds, dims, vocabs, lens, combs = self._plan_args
            plan = lib.dfm_plan_create(len(kinds), _lib.i32_array(kinds), _lib.i32_array(dims),
                                       _lib.i64_array(vocabs), _lib.i32_array(lens), _lib.i32_array(combs),
                                       int(self.fm_embed_dim))
            if not plan:
                raise ValueError(f"dfm_plan_create: {_lib.last_error()}")
            self._plan = C.c_void_p(plan)
            info = (C.c_int64 * 8)()
            _lib.check(lib.dfm_plan_info(self._plan, info), "dfm_plan_info")
            assert (info[0], info[1], info[3], bool(info[4]), info[5]) == \
                (self._T, self._S, self._A, self._aliasable, self._max_tdim), "plan/host size mismatch"
        return self._plan

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan is not None:
            try:
                _lib.lib().dfm_plan_destroy(plan)
            except Exception:
                pass

    def _apply(self, fn, *args, **kwargs):
        self._ordered_cache = None        # .to() / .cuda() may replace the Parameter objects
        self._l2_split = None
        self.__dict__.pop("_prepared_inputs", None)
        return super()._apply(fn, *args, **kwargs)

    def _ordered_params(self) -> List[torch.Tensor]:
        """Present parameters in kernel order; also records which are tables / their slot 5f+k.  Cached: the module
        tree walk costs ~0.2 ms and the step asks for the list several times."""
        cached = getattr(self, "_ordered_cache", None)
        if cached is not None:
            return cached
        out, is_table, slots = [], [], []
        for f, name in enumerate(self.field_names):
            second, first = self.second_order_embeddings[name], self.first_order_embeddings[name]
            dense = self._kinds[f] == _lib.DENSE
            entries = [(0, second.weight, not dense)]
            if dense:
                entries.append((1, second.bias, False))
            entries.append((2, first.weight, not dense))
            if dense:
                entries.append((3, first.bias, False))
            if name in self.projections:
                entries.append((4, self.projections[name].weight, False))
            for k, p, tab in entries:
                out.append(p)
                is_table.append(tab)
                slots.append(5 * f + k)
        self._param_is_table, self._slot_of_param = is_table, slots
        self._ordered_cache = out
        return out

    def _param_ptrs(self, tensors) -> C.Array:
        arr = (C.c_void_p * (5 * self.num_fields))()
        for slot, t in zip(self._slot_of_param, tensors):
            arr[slot] = None if t is None else t.data_ptr()
        return arr

    def _prepare_input(self, name: str, f: int, x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(x, f"batch[{name!r}]")
        kind = self._kinds[f]
        if kind == _lib.DENSE:
            if x.dtype != torch.float32:
                x = x.float()
            if x.dim() != 1:
                x = x.reshape(x.shape[0])
        else:
            if x.dtype != torch.int64:
                x = x.long()
            L = self._plan_args[3][f]
            if kind == _lib.SEQUENCE:
                if x.dim() != 2 or x.shape[1] != L:
                    raise ValueError(f"batch[{name!r}] must have shape (B, {L}), got {tuple(x.shape)}")
            elif x.dim() != 1:
                raise ValueError(f"batch[{name!r}] must have shape (B,), got {tuple(x.shape)}")
        return x.contiguous()

    # -- device-side input pipeline hook ------------------------------------
    def prepare(self, batch: Dict[str, torch.Tensor], stream: Optional["torch.cuda.Stream"] = None) -> None:
        """Sort a FUTURE batch's row keys now.  The backward's sort depends on the ids only, not on weights or
        gradients, so the input pipeline runs it (dfm_emit_keys + dfm_sort_keys) on a side stream while the current
        step computes; the forward of that batch then skips the key emission and its backward starts at the
        segmented reduction (DFM_GRAD_PRESORTED).  Matching is by the id tensors' storage: pass the same tensors to
        ``forward``.  Calling it is optional -- without it K1 emits the keys and K2 sorts them itself."""
        if self._S == 0:
            return
        self._ensure_plan()
        lib = _lib.lib()
        inputs = [self._prepare_input(n, f, batch[n]) for f, n in enumerate(self.field_names)]
        dev = inputs[0].device
        B = inputs[0].shape[0]
        N = B * self._S
        if N == 0:
            return
        cur = torch.cuda.current_stream(dev)
        side = stream if stream is not None else self._side_stream(dev)
        side.wait_stream(cur)                      # the id tensors are ready on the caller's stream
        with torch.cuda.stream(side):
            keys = torch.empty((N,), device=dev, dtype=torch.int32)
            _lib.check(lib.dfm_emit_keys(self._plan, B, _lib.ptr_array(inputs), keys.data_ptr(), side.cuda_stream), "dfm_emit_keys")
        prepared = self._sort_async(keys, N, dev, side=side, wait=False)
        for t in inputs:
            t.record_stream(side)
        sig = tuple(t.data_ptr() for t, k in zip(inputs, self._kinds) if k != _lib.DENSE) + (B,)
        self._prepared = (sig, prepared)

    def _side_stream(self, dev):
        side = getattr(self, "_prep_stream", None)
        if side is None or side.device != dev:
            side = self._prep_stream = torch.cuda.Stream(device=dev)
        return side

    def _sort_async(self, keys, N: int, dev, side=None, wait: bool = True):
        """dfm_sort_keys on a side stream; returns (sorted_keys, sorted_payload, event)."""
        lib = _lib.lib()
        cur = torch.cuda.current_stream(dev)
        if side is None:
            side = self._side_stream(dev)
        if wait:
            side.wait_stream(cur)
        with torch.cuda.stream(side):
            skeys = torch.empty((N,), device=dev, dtype=torch.int32)
            spay = torch.empty((N,), device=dev, dtype=torch.int32)
            ws = torch.empty((max(lib.dfm_sort_keys_workspace_bytes(self._plan, N), 16),), device=dev, dtype=torch.uint8)
            store = getattr(self, "profile_events", None)
            t0 = t1 = None
            if store is not None:
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record(side)
            _lib.check(lib.dfm_sort_keys(self._plan, N, keys.data_ptr(), skeys.data_ptr(), spay.data_ptr(), ws.data_ptr(),
                                         ws.numel(), side.cuda_stream), "dfm_sort_keys")
            if store is not None:
                t1.record(side)
                store.setdefault("sort", []).append((t0, t1))
            ev = torch.cuda.Event()
            ev.record(side)
        keys.record_stream(side)
        return skeys, spay, ev

    def _take_prepared(self, inputs):
        pref = getattr(self, "_prepared", None)
        if pref is None:
            return None
        sig = tuple(t.data_ptr() for t, k in zip(inputs, self._kinds) if k != _lib.DENSE) + (inputs[0].shape[0],)
        if pref[0] != sig:
            return None
        self._prepared = None
        return pref[1]

    def forward_fused(self, batch: Dict[str, torch.Tensor]):
        """(first_order, field_embeddings, flat, fm_value) -- the FM term comes for free."""
        self._ensure_plan()
        if self._status_pending and self.check_indices:
            self.raise_if_bad_index(block=True, keep=1)
        params = self._ordered_params()
        if not params[0].is_cuda or not params[-1].is_cuda:      # one device per module: the ends stand for all
            _lib.require_cuda(params[0], "FeatureEmbedding parameter")
            _lib.require_cuda(params[-1], "FeatureEmbedding parameter")
        # validated / normalised inputs of the last few batches, by the identity of their tensors (staging rings and
        # rotating device batches present the same tensor objects again: 39 dtype / shape checks saved per step)
        key = tuple(map(id, batch.values()))
        cache = self.__dict__.setdefault("_prepared_inputs", {})
        hit = cache.get(key)
        if hit is not None and all(a is b for a, b in zip(hit[0], batch.values())):
            inputs = hit[1]
        else:
            inputs = [self._prepare_input(n, f, batch[n]) for f, n in enumerate(self.field_names)]
            B = inputs[0].shape[0]
            for n, x in zip(self.field_names, inputs):
                if x.shape[0] != B:
                    raise ValueError(f"batch[{n!r}] has {x.shape[0]} rows, expected {B}")
            while len(cache) >= 8:
                cache.pop(next(iter(cache)))
            cache[key] = (tuple(batch.values()), inputs)       # holds the tensors: an id cannot be recycled while cached
        need_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        first, field, flat, fm, anchor = _EmbedFn.apply(self, len(inputs), need_bwd, *inputs, *params)
        self._live_anchor = anchor if (need_bwd and anchor.requires_grad) else None   # the L2 node hangs itself below this node
        field._dfm_fm = (fm, field._version)     # picked up by FMInteraction (same tensor object)
        return first, field, flat, fm

    def forward(self, batch: Dict[str, torch.Tensor]):
        first, field, flat, _ = self.forward_fused(batch)
        return first, field, flat

    # -- row-sparse helpers ------------------------------------------------------------------
    def materialize_sparse_grads(self) -> None:
        """Turn ``self.row_grads`` into ``torch.sparse_coo`` ``.grad`` tensors (one host sync),
        the layout ``nn.Embedding(sparse=True)`` produces and ``torch.optim.SparseAdam`` eats."""
        if self.row_grads is None:
            return
        for name, (rows, g2, g1) in self.row_grads.per_table().items():
            w2 = self.second_order_embeddings[name].weight
            w1 = self.first_order_embeddings[name].weight
            w2.grad = torch.sparse_coo_tensor(rows[None], g2, w2.shape).coalesce()
            w1.grad = torch.sparse_coo_tensor(rows[None], g1[:, None], w1.shape).coalesce()
