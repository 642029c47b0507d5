"""Parity ON THE SHAPES THE BENCH PUBLISHES (BASELINE.json configs).

  * Criteo-shaped schema (`workloads.criteo_schema`: all 39 fields, 13 DENSE + 26 SPARSE, D = 64, vocabularies scaled
    to 2 % so the numpy oracle fits, one table > 2^16 rows), B = 16384, log-uniform ("Zipf") ids -- the same generator
    `bench.py` uses: embedding views, FM and every gradient against the numpy oracle in `dense` mode, `row_sparse`
    against `dense` on the touched rows, the pre-sorted (input pipeline) backward bit-identical to the in-line one.
  * Whole models on the bench schemas against the torch-CPU port of the reference (`oracle/torch_port.py`, pinned to
    the reference's golden fixtures by tests/test_oracle.py): DeepFM on the Criteo shape, DeepFM / xDeepFM /
    AttentionDeepFM on `ml100k_schema()` (14 of 16 fields projected), xDeepFM on the multi-hot Criteo schema.
fp32 tolerances: logits 1e-5, gradients 1e-4 per-tensor max-norm relative (absolute floor 1e-6 for
analytically-zero gradients).
"""

import numpy as np
import pytest
import torch

from deepfm_b200 import workloads as W
from deepfm_b200.config import ExperimentConfig
from deepfm_b200.layers.embedding import FeatureEmbedding
from deepfm_b200.layers.fm import FMInteraction
from deepfm_b200.layers.l2 import l2_penalty
from deepfm_b200.models import create_model
from oracle import deepfm_oracle as O
from oracle import torch_port as TP
from tests.helpers import assert_close_rel, grads_of

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-5, 1e-4


def _cuda(batch):
    return {k: v.cuda() for k, v in batch.items()}


def _np(batch):
    return {k: v.numpy() for k, v in batch.items()}


def test_criteo_shape_embedding_fm_and_grads_vs_oracle():
    D, B = 64, 16384
    schema = W.criteo_schema(D, vocab_scale=0.02)
    assert len(schema.fields) == 39 and max(f.vocabulary_size for f in schema.fields.values()) > 2 ** 16
    torch.manual_seed(0)
    emb = FeatureEmbedding(schema, D).cuda()
    with torch.no_grad():
        for p in emb.parameters():
            p.add_(0.02 * torch.randn_like(p))        # non-zero padding rows and biases
    params = {k: v.detach().cpu().numpy() for k, v in emb.state_dict().items()}
    batch = W.synthetic_batch(schema, B, seed=3)
    rng = np.random.default_rng(1)
    g_first = rng.standard_normal((B, 1)).astype(np.float32)
    g_flat = (rng.standard_normal((B, schema.total_embedding_dim)) * 0.1).astype(np.float32)
    g_fm = (rng.standard_normal((B, 1)) * 0.05).astype(np.float32)
    lam = 1e-4

    def run(mode, prepare=False):
        emb.grad_mode = mode
        emb.zero_grad(set_to_none=True)
        dev = _cuda(batch)
        if prepare:
            emb.prepare(dev)
        fo, fe, fl = emb(dev)
        fm = FMInteraction()(fe)
        loss = (fo * torch.from_numpy(g_first).cuda()).sum() + (fl * torch.from_numpy(g_flat).cuda()).sum() \
            + (fm * torch.from_numpy(g_fm).cuda()).sum() + l2_penalty(emb, lam)
        loss.backward()
        return fo, fe, fl, fm

    fo, fe, fl, fm = run("dense")
    assert fe.data_ptr() == fl.data_ptr()                  # stack(dim=1) == cat(dim=-1): written once
    ref = O.embedding_forward(schema, params, _np(batch), D)
    assert_close_rel(fo.detach().cpu(), ref["first_order"], FWD_TOL, "first_order")
    assert np.array_equal(fl.detach().cpu().numpy()[:, 13 * D:], ref["flat"][:, 13 * D:])   # plain gathers are exact
    assert_close_rel(fl.detach().cpu(), ref["flat"], FWD_TOL, "flat")
    assert_close_rel(fm.detach().cpu(), O.fm_forward(ref["field_embeddings"].astype(np.float64)), 2e-5, "fm")
    ge = O.fm_backward(ref["field_embeddings"], g_fm)
    want = O.embedding_backward(schema, params, _np(batch), D, g_first, ge, g_flat, l2_reg=lam)
    dense = grads_of(emb)
    for k in want:
        assert_close_rel(dense[k], want[k], GRAD_TOL, k)
    # row-sparse == dense on the touched rows, bit for bit; the pre-sorted backward (input pipeline) too
    for prepare in (False, True):
        run("row_sparse", prepare)
        n_valid, n_unique = emb.last_counts.cpu().tolist()
        assert n_valid == B * 26
        per = emb.row_grads.per_table()
        assert sum(r.numel() for r, _, _ in per.values()) == n_unique
        for name, (rows, g2, g1) in per.items():
            d2 = torch.from_numpy(dense[f"second_order_embeddings.{name}.weight"]).cuda()
            d1 = torch.from_numpy(dense[f"first_order_embeddings.{name}.weight"]).cuda()
            assert torch.equal(d2[rows], g2) and torch.equal(d1[rows, 0], g1), (name, prepare)
    run("dense", prepare=True)
    again = grads_of(emb)
    for k in dense:
        assert np.array_equal(again[k], dense[k]), k       # deterministic, and identical with pre-sorted keys


def _bench_cfg(fm_dim, cin=None):
    cfg = ExperimentConfig()
    cfg.feature.fm_embed_dim = fm_dim
    cfg.dnn.dropout = 0.0              # BatchNorm keeps its training-mode batch statistics; dropout would need shared RNG
    if cin is not None:
        cfg.cin.layer_sizes = cin
    return cfg


def _port_grads(name, schema, cfg, params, batch, labels, dtype):
    cast = lambda t: t.to(dtype) if t.is_floating_point() else t
    port = TP.PortedModel(name, schema, cfg, params={k: cast(v) for k, v in params.items()})
    b = {k: cast(v) for k, v in batch.items()}
    logits = port.forward(b)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.squeeze(1), labels.to(dtype)) + port.l2_reg_loss()
    loss.backward()
    return logits.detach(), loss.detach(), {k: v.grad for k, v in port.params.items() if v.grad is not None}


def _whole_model_vs_port(name, schema, cfg, B, seed, logits_tol=1e-5, grad_tol=GRAD_TOL, cin_precision=None):
    """The CUDA model against the torch-CPU port of the reference on the same weights and batch.  Tolerance per
    tensor: the stated fp32 bound (logits 1e-5, gradients 1e-4 max-norm relative), or -- where a long fp32 reduction
    (batch-size-long sums behind BatchNorm) makes the fp32 REFERENCE itself deviate more than that from its own fp64
    evaluation -- three times that measured reference-fp32-vs-fp64 deviation (printed)."""
    torch.manual_seed(seed)
    model = create_model(name, schema, cfg).cuda().train()
    if cin_precision is not None:
        model.cin.precision = cin_precision
    params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    batch = W.synthetic_batch(schema, B, seed=seed)
    labels = W.synthetic_labels(B, seed=seed)
    logits = model(_cuda(batch))
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits.squeeze(1), labels.cuda()) + model.get_l2_reg_loss()
    loss.backward()
    l32, loss32, g32 = _port_grads(name, schema, cfg, params, batch, labels, torch.float32)
    l64, loss64, g64 = _port_grads(name, schema, cfg, params, batch, labels, torch.float64)
    from tests.conftest import rel_max_err
    noise = rel_max_err(l32.numpy(), l64.numpy(), floor=2e-6)
    assert_close_rel(logits.detach().cpu(), l64, max(logits_tol, 3 * noise), f"{name} logits", floor=2e-6)
    assert abs(loss.item() - loss64.item()) <= 1e-5 * abs(loss64.item()), (loss.item(), loss64.item())
    worst = []
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        noise = rel_max_err(g32[k].numpy(), g64[k].numpy(), floor=1e-6)
        ours = rel_max_err(p.grad.cpu().numpy(), g64[k].numpy(), floor=1e-6)
        worst.append((ours, noise, k))
        assert_close_rel(p.grad.cpu(), g64[k], max(grad_tol, 3 * noise), f"{name} {k} (reference fp32 vs fp64: {noise:.2e})", floor=1e-6)
    worst.sort(reverse=True)
    print(f"[{name}] worst gradient deviations vs fp64 (ours, reference-fp32): " +
          ", ".join(f"{k}: {o:.1e}/{n:.1e}" for o, n, k in worst[:3]))


def test_deepfm_on_criteo_shape_vs_reference_port():
    schema = W.criteo_schema(64, vocab_scale=0.01)
    _whole_model_vs_port("deepfm", schema, _bench_cfg(64), B=8192, seed=5)


@pytest.mark.parametrize("name,cin", [("deepfm", None), ("xdeepfm", [64]), ("xdeepfm", [128, 128, 64]), ("attention_deepfm", None)])
def test_models_on_ml100k_schema_vs_reference_port(name, cin):
    """BASELINE configs 1-3: the ML-100K schema (16 fields, 14 of them projected to fm_embed_dim 16, one mean bag)."""
    # CIN "auto" would pick the TF32 tensor-core path at this batch (2e-3 tolerance, pinned in test_cin_attention_gpu);
    # the whole-model fp32 bound is checked on the reference's own arithmetic.
    _whole_model_vs_port(name, W.ml100k_schema(), _bench_cfg(16, cin), B=4096, seed=7,
                         cin_precision="fp32" if cin is not None else None)


def test_xdeepfm_on_multihot_criteo_schema_vs_reference_port():
    """BASELINE config 5 shape: 13 DENSE + 26 mean-pooled bags (L = 16, ~8 ids), D = 64, CIN [128, 128] -- the
    reference materialises a (B, 39*39, 64) outer product on the CPU, hence the small batch."""
    schema = W.criteo_multihot_schema(64, max_length=16, vocab_scale=0.0005)
    _whole_model_vs_port("xdeepfm", schema, _bench_cfg(64, [128, 128]), B=384, seed=9, cin_precision="fp32")


def test_table_norm_cache_tracks_row_sparse_adam_without_dense_passes():
    """||W||^2 is maintained from the touched rows (dfm_adam_rows adds sum(new^2 - old^2)); after
    many steps it still equals the exact reduction to 1e-6 relative, and no dense pass ran after the first."""
    from deepfm_b200.layers.l2 import table_norm_cache
    from deepfm_b200.optim import RowSparseAdam
    D, B = 64, 2048
    schema = W.criteo_schema(D, vocab_scale=0.002)
    torch.manual_seed(2)
    emb = FeatureEmbedding(schema, D).cuda()
    emb.grad_mode = "row_sparse"
    opt = RowSparseAdam(emb, lr=5e-2)
    lam = 1e-3
    steps = 200
    for step in range(steps):
        batch = _cuda(W.synthetic_batch(schema, B, seed=100 + step))
        fo, fe, fl = emb(batch)
        val = l2_penalty(emb, lam)
        ((fl * torch.randn_like(fl)).sum() + fo.sum() + val).backward()
        opt.step()
    cache = table_norm_cache(emb)
    assert cache.refreshes == 1                                   # one exact pass at step 0, none afterwards
    tables = [p for p, t in zip(emb._ordered_params(), emb._param_is_table) if t]
    assert cache.valid_for(tables)
    exact = sum(float((p.detach().double() ** 2).sum().item()) for p in tables)
    got = float(cache.acc.item())
    assert abs(got - exact) <= 1e-6 * exact, (got, exact)
    small = [p for p, t in zip(emb._ordered_params(), emb._param_is_table) if not t]
    want = lam * (exact + sum(float((p.detach().double() ** 2).sum().item()) for p in small))
    assert abs(l2_penalty(emb, lam).item() - want) <= 2e-6 * want
