"""Pin the oracle: numpy restatement and torch-CPU port vs the reference's golden facts and
vs fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""

import numpy as np
import pytest
import torch

from oracle import deepfm_oracle as O
from oracle import torch_port as TP
from tests.conftest import load_golden, rel_max_err, split_prefixed
from tests.golden import spec

FP32_TOL = 2e-6   # fp32 restatement vs fp32 reference: only summation order differs


# ----------------------------------------------------------------- reference's own golden facts

def test_fm_worked_example_from_reference_notes():
    # notes/deepfm.md:72-91: v=[[1,2],[3,4],[5,6]] -> 67
    e = np.array([[[1, 2], [3, 4], [5, 6]]], dtype=np.float32)
    assert O.fm_forward(e).item() == 67.0
    g = load_golden("fm.npz")
    assert g["example_out"].item() == 67.0


def test_fm_equals_explicit_pairwise():
    # tests/test_layers.py:79-92, atol=rtol=1e-5
    rng = np.random.default_rng(0)
    e = rng.standard_normal((2, 4, 8)).astype(np.float32)
    np.testing.assert_allclose(O.fm_forward(e), O.fm_pairwise(e), atol=1e-5, rtol=1e-5)


def test_fm_single_field_is_zero():
    # tests/test_layers.py:94-98
    e = np.random.default_rng(1).standard_normal((2, 1, 8)).astype(np.float32)
    assert np.all(O.fm_forward(e) == 0)


def test_all_zero_indices_give_exactly_zero_views():
    # tests/test_layers.py:43-51 (fresh init: row 0 of every table is zero)
    schema = spec.golden_schema()
    gen = torch.Generator().manual_seed(0)
    params = {k: v.numpy() for k, v in TP.init_embedding_params(schema, spec.FM_DIM, gen).items()}
    batch = {n: np.zeros_like(v) for n, v in spec.golden_batch().items()}
    sparse_only = {k: f for k, f in schema.fields.items() if O._kind(f) != "dense"}
    sub = type(schema)(fields=sparse_only)
    out = O.embedding_forward(sub, params, batch, spec.FM_DIM)
    assert np.abs(out["first_order"]).sum() == 0
    assert np.abs(out["field_embeddings"]).sum() == 0
    assert np.abs(out["flat"]).sum() == 0


# ----------------------------------------------------------------- fixtures from the reference

def test_embedding_forward_matches_reference():
    g = load_golden("embedding.npz")
    schema = spec.golden_schema()
    out = O.embedding_forward(schema, split_prefixed(g, "param/"), split_prefixed(g, "batch/"), spec.FM_DIM)
    for k in ("first_order", "field_embeddings", "flat"):
        assert out[k].shape == g[k].shape
        assert rel_max_err(out[k], g[k]) < FP32_TOL, k


def test_embedding_backward_matches_reference_autograd():
    g = load_golden("embedding.npz")
    schema = spec.golden_schema()
    params, batch = split_prefixed(g, "param/"), split_prefixed(g, "batch/")
    grads = O.embedding_backward(schema, params, batch, spec.FM_DIM, g["g_first"], g["g_field"],
                                 g["g_flat"], l2_reg=spec.L2_REG)
    ref = split_prefixed(g, "grad/")
    assert set(grads) == set(ref)
    for k in ref:
        assert grads[k].shape == ref[k].shape, k
        assert rel_max_err(grads[k], ref[k]) < FP32_TOL, k
    assert abs(O.l2_reg_loss(params, spec.L2_REG) - float(g["l2_loss"])) < 1e-6 * float(g["l2_loss"])


def test_embedding_padding_row_semantics():
    """Row 0 is returned as stored by nn.Embedding but skipped by EmbeddingBag, and the lookup
    never sends gradient to it (SURVEY a3')."""
    g = load_golden("embedding.npz")
    params, batch = split_prefixed(g, "param/"), split_prefixed(g, "batch/")
    assert np.all(params["second_order_embeddings.u.weight"][0] == 0.25)
    b = int(np.flatnonzero(batch["u"] == 0)[0])
    assert np.all(g["flat"][b, :8] == 0.25)                     # stored row 0 comes back
    ref = split_prefixed(g, "grad/")
    np.testing.assert_allclose(ref["second_order_embeddings.u.weight"][0], 2 * spec.L2_REG * 0.25, rtol=1e-6)
    allpad = int(np.flatnonzero((batch["g"] != 0).sum(1) == 0)[0])
    off_g = 8 + 4
    assert np.all(g["flat"][allpad, off_g:off_g + 4] == 0)     # all-pad bag -> exact zero


def test_fm_matches_reference():
    g = load_golden("fm.npz")
    assert rel_max_err(O.fm_forward(g["x"]), g["out"]) < FP32_TOL
    assert rel_max_err(O.fm_backward(g["x"], g["g"]), g["gx"]) < FP32_TOL


@pytest.mark.parametrize("tag", ["split", "nosplit", "one"])
def test_cin_matches_reference(tag):
    g = load_golden(f"cin_{tag}.npz")
    n = len(g["sizes"])
    W = [g[f"w{i}"][:, :, 0] for i in range(n)]
    b = [g[f"b{i}"] for i in range(n)]
    split = bool(g["split"])
    out = O.cin_forward(g["x"], W, b, split)
    assert out.shape == g["out"].shape
    assert rel_max_err(out, g["out"]) < FP32_TOL
    gx, gW, gb = O.cin_backward(g["x"], W, b, split, g["g"])
    assert rel_max_err(gx, g["gx"]) < 5e-6
    for i in range(n):
        assert rel_max_err(gW[i], g[f"gw{i}"][:, :, 0]) < 5e-6
        assert rel_max_err(gb[i], g[f"gb{i}"]) < 5e-6


def test_cin_plan_matches_reference_split_rule():
    # cin.py:51-62: direct = L//2 first, next = L - direct; last layer keeps everything
    assert O.cin_plan(16, [128, 128, 64], True) == ([64, 64, 64], [64, 64, 64], [256, 1024, 1024])
    assert O.cin_plan(4, [6, 5, 4], True) == ([3, 2, 4], [3, 3, 4], [16, 12, 12])
    assert O.cin_plan(6, [64, 64], False) == ([64, 64], [64, 64], [36, 384])


@pytest.mark.parametrize("tag", ["res", "nores"])
def test_attention_matches_reference(tag):
    g = load_golden(f"attention_{tag}.npz")
    params, ref_grads = split_prefixed(g, "param/"), split_prefixed(g, "grad/")
    heads, res, layers = int(g["num_heads"]), bool(g["use_residual"]), int(g["num_layers"])
    xs = [g["x"]]
    for li in range(layers):
        p = split_prefixed(params, f"layers.{li}.")
        xs.append(O.attn_block_forward(xs[-1], p, heads, res))
    assert rel_max_err(xs[-1], g["out"]) < 5e-6
    gcur = g["g"]
    for li in reversed(range(layers)):
        p = split_prefixed(params, f"layers.{li}.")
        gcur, pg = O.attn_block_backward(xs[li], p, heads, res, gcur)
        for k, v in pg.items():
            ref = ref_grads[f"layers.{li}.{k}"]
            # W_k.bias grad is analytically zero (softmax shift invariance): abs tolerance only
            if k == "W_k.bias":
                assert np.abs(v - ref).max() < 1e-5
            else:
                assert rel_max_err(v, ref) < 2e-5, (li, k)
    assert rel_max_err(gcur, g["gx"]) < 2e-5


# ----------------------------------------------------------------- the torch-CPU port (cpu_baseline)

@pytest.mark.parametrize("name", ["deepfm", "xdeepfm", "attention_deepfm"])
def test_torch_port_matches_reference_model(name):
    g = load_golden(f"model_{name}.npz")
    schema, cfg = spec.golden_schema(), spec.golden_config()
    params = {k: torch.from_numpy(v) for k, v in split_prefixed(g, "param/").items()}
    model = TP.PortedModel(name, schema, cfg, params=params)
    batch = {k: torch.from_numpy(v) for k, v in spec.golden_batch().items()}
    labels = torch.from_numpy(spec.golden_labels())
    model.training = True
    logits = model.forward(batch)
    assert rel_max_err(logits.detach().numpy(), g["logits"]) < 5e-6
    # one training forward only: BatchNorm running stats must advance exactly once
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.squeeze(1), labels) + model.l2_reg_loss()
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    loss.backward()
    ref = split_prefixed(g, "grad/")
    for k, v in model.params.items():
        if k in model.buffers:
            continue
        # W_k.bias (softmax shift invariance) and a Linear bias feeding BatchNorm have
        # analytically zero gradients: pure rounding noise, so an absolute floor applies
        err = np.abs(v.grad.numpy() - ref[k]).max()
        assert err <= 1e-4 * np.abs(ref[k]).max() + 1e-6, k
    model.training = False
    with torch.no_grad():
        probs = torch.sigmoid(model.forward(batch)).numpy()
    assert rel_max_err(probs, g["probs_eval"]) < 5e-6


def test_torch_port_state_dict_keys_equal_reference():
    for name in ("deepfm", "xdeepfm", "attention_deepfm"):
        g = load_golden(f"model_{name}.npz")
        ref_keys = {k for k in split_prefixed(g, "param/") if "num_batches_tracked" not in k}
        model = TP.PortedModel(name, spec.golden_schema(), spec.golden_config())
        assert set(model.params) == ref_keys


def test_unknown_model_raises_value_error():
    with pytest.raises(ValueError):
        TP.PortedModel("unknown", spec.golden_schema(), spec.golden_config())


# ----------------------------------------------------------------- integer artefacts

def test_csr_flatten_and_counts():
    ids = spec.golden_batch()["g"]
    off, val = O.csr_flatten(ids)
    assert off.tolist() == [0, 2, 2, 5, 10, 11, 14]
    assert val.tolist() == [1, 2, 8, 8, 3, 4, 4, 4, 4, 4, 7, 2, 5, 1]
    assert O.bag_counts(ids).tolist() == [2, 0, 3, 5, 1, 3]


def test_slot_layout_keys_sort_segments():
    schema, batch = spec.golden_schema(), spec.golden_batch()
    sf, sp, rb = O.slot_layout(schema)
    assert sf.tolist() == [0, 1, 2, 2, 2, 2, 2, 4, 4, 4, 5, 5, 5, 5]
    assert rb.tolist() == [0, 11, 18, 27, 27, 33, 39, 39]
    keys, payload = O.emit_keys(schema, batch)
    assert keys.dtype == np.uint32 and len(keys) == 6 * 14
    assert keys[0] == 3 and keys[14] == O.PAD_KEY                 # u[0]=3 ; u[1]=0 -> PAD
    sk, spay = O.sort_pairs(keys, payload)
    assert np.all(sk[:-1] <= sk[1:])
    uk, starts = O.segment_heads(sk)
    assert len(uk) == len(np.unique(keys[keys != O.PAD_KEY]))
    assert starts[-1] == (keys != O.PAD_KEY).sum()
    # stability: payloads inside a segment are ascending (fixed summation order)
    for a, b in zip(starts[:-1], starts[1:]):
        assert np.all(np.diff(spay[a:b].astype(np.int64)) > 0)


def test_shard_route():
    ids = np.array([5, 8, 1, 0, 13, 2, 7, 16], dtype=np.int64)
    owner, local, counts, offsets, perm = O.shard_route(ids, 4)
    assert owner.tolist() == [1, 0, 1, 0, 1, 2, 3, 0]
    assert local.tolist() == [1, 2, 0, 0, 3, 0, 1, 4]
    assert counts.tolist() == [3, 3, 1, 1] and offsets.tolist() == [0, 3, 6, 7, 8]
    assert perm.tolist() == [1, 3, 7, 0, 2, 4, 5, 6]
    assert np.all(local * 4 + owner == ids)


def test_adam_rows_equals_torch_adam_when_every_row_is_touched():
    """Pins oracle.adam_rows (the optimiser row) to torch.optim.Adam (trainer.py:67-78 defaults):
    with every row touched at every step the row-restricted update IS dense Adam."""
    import torch
    rng = np.random.default_rng(5)
    w = rng.standard_normal((7, 4)).astype(np.float32)
    p = torch.nn.Parameter(torch.from_numpy(w.copy()))
    opt = torch.optim.Adam([p], lr=1e-3)
    m, v = np.zeros_like(w), np.zeros_like(w)
    rows = np.arange(7)
    for step in (1, 2, 3, 4):
        g = rng.standard_normal((7, 4)).astype(np.float32)
        p.grad = torch.from_numpy(g.copy())
        opt.step()
        w, m, v = O.adam_rows(w, m, v, rows, g, step)
        np.testing.assert_allclose(w, p.detach().numpy(), rtol=2e-6, atol=1e-7)
    # lazy semantics: a row that is not touched keeps weight and moments
    w2, m2, v2 = O.adam_rows(w, m, v, np.array([1, 3]), g[[1, 3]], 5)
    assert np.array_equal(w2[[0, 2, 4, 5, 6]], w[[0, 2, 4, 5, 6]]) and np.array_equal(m2[0], m[0])
