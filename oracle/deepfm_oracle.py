"""CPU oracle for the DeepFM embedding-plus-interaction hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference's algorithm, used solely as the checker
in ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` leg.  Nothing
under ``deepfm_b200/`` imports it; the product path fails loudly if the CUDA library is
missing rather than falling back to this.

Where the arithmetic lives: the reference is pure PyTorch; every op on the path executes in
the un-vendored third-party ``torch`` (ATen CPU kernels; ``uv.lock`` pins torch 2.10.0, this
image has 2.11.0).  The functions below restate the *published semantics* of those ATen ops
(``embedding`` / ``embedding_bag`` with ``padding_idx``, ``linear``, ``conv1d(k=1)``,
``softmax``, ``layer_norm``) at the reference's own call sites, each citing the reference
file:line it follows, plus hand-derived backward passes (the reference relies on autograd).

Pinning: checked in ``tests/test_oracle.py`` against
  * the reference's own golden facts for the path: FM worked example ``[[1,2],[3,4],[5,6]] -> 67``
    (notes/deepfm.md:72-91), FM == explicit pairwise sum at 1e-5 (tests/test_layers.py:79-92),
    single field -> 0 (tests/test_layers.py:94-98), all-index-0 batch -> three exactly-zero
    views (tests/test_layers.py:43-51);
  * outputs AND gradients of the unmodified reference modules imported in the build container
    (fixtures in ``tests/golden/*.npz`` written by ``tests/golden/make_golden.py``).

All functions take/return numpy arrays.  ``dt`` selects float32 (bit-comparable to the
reference up to summation order) or float64 (the error yardstick).
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

SPARSE, SEQUENCE, DENSE = "sparse", "sequence", "dense"


def _kind(field) -> str:
    ft = field.feature_type
    return ft.value if hasattr(ft, "value") else str(ft)


# --------------------------------------------------------------------------------------
# FeatureEmbedding  (reference: deepfm/models/layers/embedding.py:76-126)
# --------------------------------------------------------------------------------------

def bag_counts(ids: np.ndarray) -> np.ndarray:
    """Non-pad entries per bag of a zero-padded (B, L) id matrix.

    EmbeddingBag(padding_idx=0) skips id 0 wherever it occurs, not only at the tail
    (embedding.py:41-50; SURVEY a3' (ii)).
    """
    return (ids != 0).sum(axis=1).astype(np.int64)


def csr_flatten(ids: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(B, L) zero-padded ids -> CSR (offsets (B+1,), values (nnz,)) in row-major order.

    offsets = exclusive prefix sum of the non-pad counts; values keep the left-to-right
    order of the surviving ids.  Integer artefact: compared bit-exactly with the device.
    """
    cnt = bag_counts(ids)
    offsets = np.zeros(ids.shape[0] + 1, dtype=np.int64)
    np.cumsum(cnt, out=offsets[1:])
    values = ids[ids != 0].astype(np.int64)  # boolean mask walks row-major
    return offsets, values


def pool_bag(weight: np.ndarray, ids: np.ndarray, combiner: str):
    """EmbeddingBag(mode=combiner, padding_idx=0) on (B, L) ids  (embedding.py:41-50,91-94).

    Returns (pooled (B, d), argmax (B, d) int64 or None).  All-pad bag -> exact zeros for
    every combiner; duplicates count once per occurrence; ``max`` ties pick the first.
    """
    B, L = ids.shape
    d = weight.shape[1]
    rows = weight[ids]                                  # (B, L, d)
    valid = (ids != 0)[:, :, None]                      # (B, L, 1)
    cnt = valid.sum(axis=1)                             # (B, 1)
    if combiner in ("sum", "mean"):
        out = np.where(valid, rows, 0).astype(weight.dtype)
        acc = np.zeros((B, d), dtype=weight.dtype)
        for l in range(L):                              # left-to-right, like the kernel
            acc = acc + out[:, l, :]
        if combiner == "mean":
            acc = np.where(cnt > 0, acc / np.maximum(cnt, 1).astype(weight.dtype), 0)
        return acc.astype(weight.dtype), None
    if combiner == "max":
        masked = np.where(valid, rows, -np.inf)
        arg = masked.argmax(axis=1)                     # first max
        out = np.take_along_axis(masked, arg[:, None, :], axis=1)[:, 0, :]
        out = np.where(cnt > 0, out, 0).astype(weight.dtype)
        return out, arg.astype(np.int64)
    raise ValueError(f"unknown combiner {combiner!r}")


def embedding_forward(schema, params: Dict[str, np.ndarray], batch: Dict[str, np.ndarray],
                      fm_embed_dim: int, dt=np.float32):
    """The three views of FeatureEmbedding.forward (embedding.py:76-126).

    ``params`` uses the reference's state_dict keys (``second_order_embeddings.<f>.weight`` ...).
    Returns dict(first_order (B,1), field_embeddings (B,F,D), flat (B,T), raw list, aux).
    nn.Embedding(padding_idx=0) returns weight[0] *as stored* (SURVEY a3' (i)); EmbeddingBag
    skips id 0.
    """
    names = list(schema.fields.keys())
    B = len(next(iter(batch.values())))
    fo_sum = np.zeros((B, 1), dtype=dt)
    raws, projs, aux = [], [], {}
    for name in names:
        f = schema.fields[name]
        k = _kind(f)
        x = np.asarray(batch[name])
        w2 = params[f"second_order_embeddings.{name}.weight"].astype(dt)
        w1 = params[f"first_order_embeddings.{name}.weight"].astype(dt)
        if k == DENSE:
            # nn.Linear(1, d): y = x * W[:,0] + b   (embedding.py:51-56,88-90)
            xv = x.astype(dt)[:, None]
            raw = xv * w2[:, 0][None, :] + params[f"second_order_embeddings.{name}.bias"].astype(dt)[None, :]
            fo = xv * w1[:, 0][None, :] + params[f"first_order_embeddings.{name}.bias"].astype(dt)[None, :]
        elif k == SEQUENCE:
            raw, arg2 = pool_bag(w2, x, f.combiner)
            fo, arg1 = pool_bag(w1, x, f.combiner)
            aux[name] = dict(count=bag_counts(x), argmax2=arg2, argmax1=arg1)
        else:
            raw = w2[x]                                 # embedding.py:95-98
            fo = w1[x]
        fo_sum = fo_sum + fo.astype(dt)                 # stack(...).sum(1)  embedding.py:118
        raws.append(raw.astype(dt))
        pkey = f"projections.{name}.weight"
        if pkey in params:                              # Linear(d_f, D, bias=False)  embedding.py:59-62,112
            projs.append(raw.astype(dt) @ params[pkey].astype(dt).T)
        else:
            projs.append(raw.astype(dt))
    field_emb = np.stack(projs, axis=1)                 # embedding.py:121
    flat = np.concatenate(raws, axis=1)                 # embedding.py:124
    return dict(first_order=fo_sum, field_embeddings=field_emb, flat=flat, raw=raws, aux=aux)


def embedding_backward(schema, params, batch, fm_embed_dim, g_first, g_field, g_flat,
                       l2_reg: float = 0.0, dt=np.float32) -> Dict[str, np.ndarray]:
    """Dense gradients of every FeatureEmbedding parameter, reference semantics.

    Inputs are dL/d(first_order) (B,1), dL/d(field_embeddings) (B,F,D), dL/d(flat) (B,T).
    Adds the L2 term of BaseCTRModel.get_l2_reg_loss (base.py:78-83): ``2*l2_reg*p`` on EVERY
    element of EVERY embedding parameter -- except that autograd forces the grad row 0 of
    nn.Embedding(padding_idx=0) / EmbeddingBag(padding_idx=0) *from the lookup* to zero; the
    L2 term still reaches row 0 (it is a plain norm over the parameter).
    """
    names = list(schema.fields.keys())
    fwd = embedding_forward(schema, params, batch, fm_embed_dim, dt)
    grads: Dict[str, np.ndarray] = {}
    off = 0
    for fi, name in enumerate(names):
        f = schema.fields[name]
        k = _kind(f)
        d = f.embedding_dim
        x = np.asarray(batch[name])
        w2key, w1key = f"second_order_embeddings.{name}.weight", f"first_order_embeddings.{name}.weight"
        w2 = params[w2key].astype(dt)
        w1 = params[w1key].astype(dt)
        g_raw = g_flat[:, off:off + d].astype(dt).copy()
        off += d
        pkey = f"projections.{name}.weight"
        ge = g_field[:, fi, :].astype(dt)
        if pkey in params:
            P = params[pkey].astype(dt)                 # (D, d)
            g_raw += ge @ P
            grads[pkey] = ge.T @ fwd["raw"][fi]         # (D, d)
        else:
            g_raw += ge
        gfo = g_first[:, 0].astype(dt)
        if k == DENSE:
            xv = x.astype(dt)
            grads[w2key] = (g_raw * xv[:, None]).sum(axis=0)[:, None]
            grads[f"second_order_embeddings.{name}.bias"] = g_raw.sum(axis=0)
            grads[w1key] = np.array([[(gfo * xv).sum()]], dtype=dt)
            grads[f"first_order_embeddings.{name}.bias"] = np.array([gfo.sum()], dtype=dt)
        elif k == SPARSE:
            gw2 = np.zeros_like(w2)
            gw1 = np.zeros_like(w1)
            keep = x != 0                               # padding_idx=0: no grad to row 0
            np.add.at(gw2, x[keep], g_raw[keep])
            np.add.at(gw1[:, 0], x[keep], gfo[keep])
            grads[w2key], grads[w1key] = gw2, gw1
        else:
            gw2 = np.zeros_like(w2)
            gw1 = np.zeros_like(w1)
            cnt = fwd["aux"][name]["count"]
            B, L = x.shape
            if f.combiner in ("sum", "mean"):
                scale = np.ones(B, dtype=dt) if f.combiner == "sum" else \
                    np.where(cnt > 0, 1.0 / np.maximum(cnt, 1), 0).astype(dt)
                for l in range(L):
                    keep = x[:, l] != 0
                    np.add.at(gw2, x[keep, l], g_raw[keep] * scale[keep, None])
                    np.add.at(gw1[:, 0], x[keep, l], gfo[keep] * scale[keep])
            else:  # max: gradient goes to the (first) arg-max entry per output dim
                a2 = fwd["aux"][name]["argmax2"]
                a1 = fwd["aux"][name]["argmax1"]
                for b in range(B):
                    if cnt[b] == 0:
                        continue
                    for j in range(d):
                        gw2[x[b, a2[b, j]], j] += g_raw[b, j]
                    gw1[x[b, a1[b, 0]], 0] += gfo[b]
            grads[w2key], grads[w1key] = gw2, gw1
    if l2_reg:
        for key in list(grads.keys()):
            grads[key] = grads[key] + (2.0 * l2_reg) * params[key].astype(dt)
    return grads


def l2_reg_loss(embedding_params: Dict[str, np.ndarray], l2_reg: float, dt=np.float64) -> float:
    """BaseCTRModel.get_l2_reg_loss: l2_reg * sum_p ||p||_2^2 (base.py:78-83)."""
    return float(l2_reg * sum((p.astype(dt) ** 2).sum() for p in embedding_params.values()))


# --------------------------------------------------------------------------------------
# FMInteraction  (reference: deepfm/models/layers/fm.py:18-23)
# --------------------------------------------------------------------------------------

def fm_forward(e: np.ndarray) -> np.ndarray:
    """0.5 * sum_d[(sum_f e)^2 - sum_f e^2]  -> (B, 1)."""
    s = e.sum(axis=1)
    return (0.5 * (s * s - (e * e).sum(axis=1)).sum(axis=1, keepdims=True)).astype(e.dtype)


def fm_backward(e: np.ndarray, g: np.ndarray) -> np.ndarray:
    """d/de[b,f,d] = g[b] * (S[b,d] - e[b,f,d])."""
    s = e.sum(axis=1, keepdims=True)
    return (g[:, :, None] * (s - e)).astype(e.dtype)


def fm_pairwise(e: np.ndarray) -> np.ndarray:
    """Explicit sum_{i<j} <e_i, e_j> (the identity tests/test_layers.py:79-92 pins)."""
    B, F, _ = e.shape
    out = np.zeros((B, 1), dtype=e.dtype)
    for i in range(F):
        for j in range(i + 1, F):
            out[:, 0] += (e[:, i] * e[:, j]).sum(axis=1)
    return out


# --------------------------------------------------------------------------------------
# CIN  (reference: deepfm/models/layers/cin.py:26-105)
# --------------------------------------------------------------------------------------

def cin_plan(num_fields: int, layer_sizes: Sequence[int], split_half: bool):
    """direct/next sizes and per-layer K exactly as CIN.__init__ builds them (cin.py:41-64)."""
    prev, direct, nxt, ks = num_fields, [], [], []
    for i, ls in enumerate(layer_sizes):
        ks.append(prev * num_fields)
        if split_half and i < len(layer_sizes) - 1:
            dsz = ls // 2
            direct.append(dsz)
            nxt.append(ls - dsz)
            prev = ls - dsz
        else:
            direct.append(ls)
            nxt.append(ls)
            prev = ls
    return direct, nxt, ks


def cin_forward(x0: np.ndarray, weights: List[np.ndarray], biases: List[np.ndarray],
                split_half: bool, keep: bool = False, masks: Optional[List[np.ndarray]] = None):
    """CIN.forward (cin.py:66-105).  weights[i]: (L_i, K_i) (Conv1d weight squeezed), K index = h*F + f.

    Split is ``[direct first, next second]`` along channels (cin.py:93-96).
    masks (test aid): per layer a boolean (B, L_i, D) array that REPLACES the ReLU decision ``pre > 0`` -- the
    decisions a reduced-precision implementation took -- so that its gradients can be compared entry by entry
    (a pre-activation within rounding distance of zero may legitimately land on either side).
    """
    B, F, D = x0.shape
    n = len(weights)
    direct, nxt, _ = cin_plan(F, [w.shape[0] for w in weights], split_half)
    hidden = x0
    parts, saved = [], []
    for i in range(n):
        H = hidden.shape[1]
        z = (hidden[:, :, None, :] * x0[:, None, :, :]).reshape(B, H * F, D)     # cin.py:84-87
        pre = np.einsum("lk,bkd->bld", weights[i], z) + biases[i][None, :, None]  # conv k=1
        act = np.maximum(pre, 0) if masks is None else pre * masks[i]              # cin.py:91
        if keep:
            saved.append((hidden, act))
        if split_half and i < n - 1:
            d_part, hidden = act[:, :direct[i]], act[:, direct[i]:]
        else:
            d_part, hidden = act, act
        parts.append(d_part.sum(axis=2))                                          # cin.py:102
    out = np.concatenate(parts, axis=1).astype(x0.dtype)
    return (out, saved) if keep else out


def cin_relu_margin(x0, weights, biases, split_half) -> float:
    """Smallest |pre-activation| over all layers.  A pre-activation closer to zero than the fp32 rounding of
    the contraction can land on the other side of the ReLU (cin.py:91) in an fp32 implementation, which changes
    its gradient by a whole term; tests draw inputs whose margin is clear of that."""
    B, F, D = x0.shape
    n = len(weights)
    direct, _, _ = cin_plan(F, [w.shape[0] for w in weights], split_half)
    hidden, margin = x0, np.inf
    for i in range(n):
        H = hidden.shape[1]
        z = (hidden[:, :, None, :] * x0[:, None, :, :]).reshape(B, H * F, D)
        pre = np.einsum("lk,bkd->bld", weights[i], z) + biases[i][None, :, None]
        margin = min(margin, float(np.abs(pre).min()))
        act = np.maximum(pre, 0)
        hidden = act[:, direct[i]:] if (split_half and i < n - 1) else act
    return margin


def cin_backward(x0, weights, biases, split_half, g_out, masks: Optional[List[np.ndarray]] = None):
    """Gradients of CIN.forward w.r.t. x0, every conv weight and bias.

    g_out: (B, output_dim).  Z is recomputed, never stored.  masks: see cin_forward.
    """
    B, F, D = x0.shape
    n = len(weights)
    direct, nxt, _ = cin_plan(F, [w.shape[0] for w in weights], split_half)
    _, saved = cin_forward(x0, weights, biases, split_half, keep=True, masks=masks)
    gx0 = np.zeros_like(x0)
    gW = [None] * n
    gb = [None] * n
    col_off = np.concatenate([[0], np.cumsum(direct)])
    g_hidden_next = None
    for i in reversed(range(n)):
        hidden, act = saved[i]
        L = weights[i].shape[0]
        g_act = np.zeros((B, L, D), dtype=x0.dtype)
        gd = g_out[:, col_off[i]:col_off[i + 1]][:, :, None]        # sum over D -> broadcast
        if split_half and i < n - 1:
            g_act[:, :direct[i]] += gd
            g_act[:, direct[i]:] += g_hidden_next
        else:
            g_act += gd
            if g_hidden_next is not None:
                g_act += g_hidden_next
        g_pre = g_act * ((act > 0) if masks is None else masks[i])
        H = hidden.shape[1]
        z = (hidden[:, :, None, :] * x0[:, None, :, :]).reshape(B, H * F, D)
        gW[i] = np.einsum("bld,bkd->lk", g_pre, z)
        gb[i] = g_pre.sum(axis=(0, 2))
        gz = np.einsum("lk,bld->bkd", weights[i], g_pre).reshape(B, H, F, D)
        g_hidden = (gz * x0[:, None, :, :]).sum(axis=2)
        gx0 += (gz * hidden[:, :, None, :]).sum(axis=1)
        if i == 0:
            gx0 += g_hidden                                          # hidden_0 is x0 itself
        else:
            g_hidden_next = g_hidden
    return gx0, gW, gb


# --------------------------------------------------------------------------------------
# Field self-attention block  (reference: deepfm/models/layers/attention.py:91-120)
# --------------------------------------------------------------------------------------

def _softmax(s):
    m = s.max(axis=-1, keepdims=True)
    e = np.exp(s - m)
    return e / e.sum(axis=-1, keepdims=True)


def attn_block_forward(x, p: Dict[str, np.ndarray], num_heads: int, use_residual: bool,
                       eps: float = 1e-5, keep: bool = False):
    """One _AttentionBlock.  p keys: W_q/W_k/W_v/W_out .weight/.bias, layer_norm.weight/.bias."""
    B, F, D = x.shape
    A = p["W_q.weight"].shape[0]
    hd = A // num_heads
    q = x @ p["W_q.weight"].T + p["W_q.bias"]
    k = x @ p["W_k.weight"].T + p["W_k.bias"]
    v = x @ p["W_v.weight"].T + p["W_v.bias"]
    qh = q.reshape(B, F, num_heads, hd).transpose(0, 2, 1, 3)
    kh = k.reshape(B, F, num_heads, hd).transpose(0, 2, 1, 3)
    vh = v.reshape(B, F, num_heads, hd).transpose(0, 2, 1, 3)
    s = qh @ kh.transpose(0, 1, 3, 2) / math.sqrt(hd)               # attention.py:80,105
    pr = _softmax(s)
    o = (pr @ vh).transpose(0, 2, 1, 3).reshape(B, F, A)            # attention.py:109-112
    y = o @ p["W_out.weight"].T + p["W_out.bias"]
    cache = dict(q=qh, k=kh, v=vh, pr=pr, o=o, y=y)
    if use_residual:
        r = y + x
        mu = r.mean(axis=-1, keepdims=True)
        var = ((r - mu) ** 2).mean(axis=-1, keepdims=True)          # biased, like LayerNorm
        rstd = 1.0 / np.sqrt(var + eps)
        xhat = (r - mu) * rstd
        out = xhat * p["layer_norm.weight"] + p["layer_norm.bias"]
        cache.update(xhat=xhat, rstd=rstd)
    else:
        out = y
    out = out.astype(x.dtype)
    return (out, cache) if keep else out


def attn_block_backward(x, p, num_heads, use_residual, g_out, eps: float = 1e-5):
    """Gradients of one block w.r.t. x and every parameter (hand-derived)."""
    B, F, D = x.shape
    A = p["W_q.weight"].shape[0]
    hd = A // num_heads
    _, c = attn_block_forward(x, p, num_heads, use_residual, eps, keep=True)
    g = {}
    if use_residual:
        gam = p["layer_norm.weight"]
        g["layer_norm.weight"] = (g_out * c["xhat"]).sum(axis=(0, 1))
        g["layer_norm.bias"] = g_out.sum(axis=(0, 1))
        gxh = g_out * gam
        gr = c["rstd"] * (gxh - gxh.mean(axis=-1, keepdims=True)
                          - c["xhat"] * (gxh * c["xhat"]).mean(axis=-1, keepdims=True))
        gy, gx = gr, gr.copy()
    else:
        gy, gx = g_out, np.zeros_like(x)
    g["W_out.weight"] = np.einsum("bfd,bfa->da", gy, c["o"])
    g["W_out.bias"] = gy.sum(axis=(0, 1))
    go = (gy @ p["W_out.weight"]).reshape(B, F, num_heads, hd).transpose(0, 2, 1, 3)
    gpr = go @ c["v"].transpose(0, 1, 3, 2)
    gv = c["pr"].transpose(0, 1, 3, 2) @ go
    gs = c["pr"] * (gpr - (gpr * c["pr"]).sum(axis=-1, keepdims=True)) / math.sqrt(hd)
    gq = gs @ c["k"]
    gk = gs.transpose(0, 1, 3, 2) @ c["q"]
    merge = lambda t: t.transpose(0, 2, 1, 3).reshape(B, F, A)
    for nm, gt in (("W_q", merge(gq)), ("W_k", merge(gk)), ("W_v", merge(gv))):
        g[f"{nm}.weight"] = np.einsum("bfa,bfd->ad", gt, x)
        g[f"{nm}.bias"] = gt.sum(axis=(0, 1))
        gx = gx + gt @ p[f"{nm}.weight"]
    return gx.astype(x.dtype), g


# --------------------------------------------------------------------------------------
# Integer artefacts of the sparse backward and of row sharding (bit-exact contracts)
# --------------------------------------------------------------------------------------

def slot_layout(schema):
    """Id-slot enumeration used by the sorted-index backward.

    Every SPARSE field owns 1 slot per sample and every SEQUENCE field ``max_length`` slots;
    DENSE fields own none.  Returns (slot_field (S,), slot_pos (S,), row_base (F+1,)) where
    row_base is the exclusive prefix sum of vocabulary sizes over id fields in schema order
    (dense fields contribute 0 rows): global_row = row_base[f] + id.
    """
    slot_field, slot_pos, row_base = [], [], [0]
    for fi, f in enumerate(schema.fields.values()):
        k = _kind(f)
        if k == SPARSE:
            slot_field.append(fi); slot_pos.append(0)
        elif k == SEQUENCE:
            for l in range(f.max_length):
                slot_field.append(fi); slot_pos.append(l)
        row_base.append(row_base[-1] + (f.vocabulary_size if k != DENSE else 0))
    return (np.asarray(slot_field, np.int32), np.asarray(slot_pos, np.int32),
            np.asarray(row_base, np.int64))


PAD_KEY = np.uint32(0xFFFFFFFF)


def emit_keys(schema, batch) -> Tuple[np.ndarray, np.ndarray]:
    """(keys uint32 (B*S,), payload uint32 (B*S,)): key = global row of the id in that slot,
    PAD_KEY for id 0; payload = (b << ceil(log2 S)) | slot  (sample and slot, shift/mask decodable)."""
    slot_field, slot_pos, row_base = slot_layout(schema)
    names = list(schema.fields.keys())
    B = len(next(iter(batch.values())))
    S = len(slot_field)
    keys = np.empty((B, S), dtype=np.uint32)
    for s in range(S):
        f = schema.fields[names[slot_field[s]]]
        x = np.asarray(batch[names[slot_field[s]]])
        ids = x if x.ndim == 1 else x[:, slot_pos[s]]
        keys[:, s] = np.where(ids != 0, row_base[slot_field[s]] + ids, PAD_KEY).astype(np.uint32)
    bits = int(np.ceil(np.log2(S))) if S > 1 else 0
    payload = ((np.arange(B, dtype=np.uint32)[:, None] << np.uint32(bits)) | np.arange(S, dtype=np.uint32)[None, :])
    return keys.reshape(-1), payload.reshape(-1).astype(np.uint32)


def sort_pairs(keys: np.ndarray, payload: np.ndarray):
    """Stable ascending sort by key (what an LSD radix sort produces)."""
    order = np.argsort(keys, kind="stable")
    return keys[order], payload[order]


def segment_heads(sorted_keys: np.ndarray):
    """(unique_keys, segment_start (U+1,)) over the non-PAD prefix of a sorted key array."""
    n_valid = int(np.searchsorted(sorted_keys, PAD_KEY, side="left"))
    k = sorted_keys[:n_valid]
    if n_valid == 0:
        return k[:0], np.zeros(1, dtype=np.int64)
    head = np.concatenate([[True], k[1:] != k[:-1]])
    starts = np.flatnonzero(head).astype(np.int64)
    return k[head], np.concatenate([starts, [n_valid]])


def shard_route(ids: np.ndarray, world: int, sent: Optional[np.ndarray] = None, rot: Optional[np.ndarray] = None):
    """Row sharding rule: owner = (id + rot) mod W, local_row = id div W  (SURVEY 8(e)), where rot is the
    schema index of the slot's field (0 when not given): a plain ``id mod W`` would put the hottest ids
    (1, 2, ... under Zipf-like skew) of EVERY table on the same ranks; rotating by the field index spreads
    them, and ``id div W`` stays a dense local row index because the ids a rank owns in one table are
    still one residue class.

    ids: flat int64 array in source order (slot index b*S + s, multi-hot bags expanded to their
    max_length slots).  sent: optional bool mask; False marks slots that are not exchanged -- the
    padding entries (id 0) of EmbeddingBag fields, which the reference skips wherever they occur
    (embedding.py:41-50, padding_idx=0).  Returns (owner, local_row, counts (W,), offsets (W+1,), perm)
    where perm is the stable permutation that groups the SENT slots by owner (position i of the send
    buffer holds ids[perm[i]]).
    """
    owner = ((ids + (0 if rot is None else rot)) % world).astype(np.int64)
    local = (ids // world).astype(np.int64)
    if sent is None:
        sent = np.ones(ids.shape, dtype=bool)
    key = np.where(sent, owner, world)
    counts = np.bincount(key, minlength=world + 1).astype(np.int64)[:world]
    offsets = np.zeros(world + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    perm = np.argsort(key, kind="stable").astype(np.int64)[: int(sent.sum())]
    return owner, local, counts, offsets, perm


def shard_route_unique(ids: np.ndarray, vbase: np.ndarray, world: int, lbits: int, sent: Optional[np.ndarray] = None,
                       rot: Optional[np.ndarray] = None):
    """Unique-row exchange (the routing ShardedFeatureEmbedding uses): every DISTINCT (owner, row) of the batch travels
    once.  ids / vbase / rot / sent: flat arrays in source order (slot index b*S + s).  owner = (id + rot) mod W,
    owner-local key = vbase + id div W (vbase: rows of the sharded tables before the slot's field, every shard sized
    ceil(V / W), so the key is the same number on every rank), composite = owner << lbits | key.  The send order is the
    ascending list of distinct composites of the SENT slots (hence grouped by owner).
    Returns (send_keys (U,) = key part in send order, counts (W,), position (n,) = 1 + index of the slot's composite in
    the send order, 0 for unsent slots, composite (n,))."""
    ids = ids.astype(np.int64)
    owner = (ids + (0 if rot is None else rot)) % world
    comp = (owner << lbits) | (vbase.astype(np.int64) + ids // world)
    if sent is None:
        sent = np.ones(ids.shape, dtype=bool)
    uniq = np.unique(comp[sent])
    counts = np.bincount(uniq >> lbits, minlength=world).astype(np.int64)[:world]
    pos = np.zeros(ids.shape, dtype=np.int64)
    pos[sent] = np.searchsorted(uniq, comp[sent]) + 1
    return (uniq & ((1 << lbits) - 1)).astype(np.int64), counts, pos, comp


def shard_positions(perm: np.ndarray, b: int, lens: Sequence[int]) -> np.ndarray:
    """1-based send position of every id slot in the field-major layout the kernels use: field f's
    (b, L_f) block starts at b * slot_base[f]; 0 = not sent."""
    S = int(sum(lens))
    pos = np.zeros(b * S, dtype=np.int64)
    pos[perm] = np.arange(1, perm.size + 1)
    pos = pos.reshape(b, S)
    out, s0 = [], 0
    for L in lens:
        out.append(pos[:, s0:s0 + L].reshape(-1))
        s0 += L
    return np.concatenate(out)


def shard_positions_from(pos_flat: np.ndarray, b: int, lens: Sequence[int]) -> np.ndarray:
    """Per-slot values in source order (b*S + s) -> the field-major layout the kernels use (see shard_positions)."""
    S = int(sum(lens))
    pos = np.asarray(pos_flat).reshape(b, S)
    out, s0 = [], 0
    for L in lens:
        out.append(pos[:, s0:s0 + L].reshape(-1))
        s0 += L
    return np.concatenate(out)


# --------------------------------------------------------------------------------------
# Row-sparse Adam (the optimiser row of SURVEY 8(f)1): torch.optim.Adam's update restricted
# to touched rows ("lazy" state), reference hyper-parameters trainer.py:67-78.
# --------------------------------------------------------------------------------------

def adam_rows(w, m, v, rows, grad_rows, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, clip_scale=1.0):
    w, m, v = w.copy(), m.copy(), v.copy()
    g = grad_rows * clip_scale
    m[rows] = b1 * m[rows] + (1 - b1) * g
    v[rows] = b2 * v[rows] + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    w[rows] = w[rows] - (lr / bc1) * m[rows] / (np.sqrt(v[rows]) / math.sqrt(bc2) + eps)
    return w, m, v
