"""Row-sparse Adam against the numpy oracle ``adam_rows`` (torch.optim.Adam's update,
trainer.py:67-78, restricted to the touched rows) and against torch.optim.Adam itself on the touched rows;
the tables' share of the global gradient norm against numpy.  fp32: 1e-6 relative on the updated weights."""

import numpy as np
import pytest
import torch

from deepfm_b200.layers.embedding import FeatureEmbedding
from deepfm_b200.layers.fm import FMInteraction
from deepfm_b200.optim import RowSparseAdam
from deepfm_b200.schema import DatasetSchema, FeatureType, FieldSchema
from oracle import deepfm_oracle as O
from tests.helpers import assert_close_rel

pytestmark = pytest.mark.gpu


def _schema(D):
    fields = {"d0": FieldSchema("d0", FeatureType.DENSE, embedding_dim=D)}
    for i, v in enumerate((50, 3000, 7)):
        fields[f"s{i}"] = FieldSchema(f"s{i}", FeatureType.SPARSE, vocabulary_size=v, embedding_dim=D)
    return DatasetSchema(fields=fields)


@pytest.mark.parametrize("D", [16, 64])
def test_row_sparse_adam_matches_oracle_and_torch_adam(D):
    torch.manual_seed(0)
    rng = np.random.default_rng(D)
    schema = _schema(D)
    B = 257
    emb = FeatureEmbedding(schema, D).cuda()
    emb.grad_mode = "row_sparse"
    opt = RowSparseAdam(emb, lr=1e-2)
    names = [n for n, f in schema.fields.items() if f.feature_type != FeatureType.DENSE]
    w2 = {n: emb.second_order_embeddings[n].weight.detach().cpu().numpy().copy() for n in names}
    w1 = {n: emb.first_order_embeddings[n].weight.detach().cpu().numpy().copy() for n in names}
    m2 = {n: np.zeros_like(w2[n]) for n in names}; v2 = {n: np.zeros_like(w2[n]) for n in names}
    m1 = {n: np.zeros_like(w1[n]) for n in names}; v1 = {n: np.zeros_like(w1[n]) for n in names}
    fm = FMInteraction()
    for step in (1, 2, 3):
        batch = {"d0": torch.from_numpy(rng.uniform(-1, 1, B).astype(np.float32)).cuda()}
        for n in names:
            V = schema.fields[n].vocabulary_size
            batch[n] = torch.from_numpy(((rng.zipf(1.4, B) - 1) % V).astype(np.int64)).cuda()
        emb.zero_grad(set_to_none=True)
        fo, fe, fl = emb(batch)
        ((fl * torch.randn_like(fl)).sum() + fo.sum() + fm(fe).sum()).backward()
        per = {n: tuple(t.cpu().numpy() for t in v) for n, v in emb.row_grads.per_table().items()}
        # tables' share of the squared gradient norm
        want_ss = sum(float((g2.astype(np.float64) ** 2).sum() + (g1.astype(np.float64) ** 2).sum()) for _, g2, g1 in per.values())
        got_ss = float(opt.grad_sumsq().item())
        assert abs(got_ss - want_ss) <= 1e-5 * want_ss
        clip = min(1.0, 1.0 / (np.sqrt(want_ss) + 1e-6))           # clip_grad_norm_(max_norm=1.0) coefficient
        opt.step(torch.tensor([clip], device="cuda", dtype=torch.float32))
        for n in names:
            rows, g2, g1 = per[n]
            w2[n], m2[n], v2[n] = O.adam_rows(w2[n], m2[n], v2[n], rows, g2, step, lr=1e-2, clip_scale=np.float32(clip))
            w1[n], m1[n], v1[n] = O.adam_rows(w1[n], m1[n], v1[n], rows, g1[:, None], step, lr=1e-2, clip_scale=np.float32(clip))
            assert_close_rel(emb.second_order_embeddings[n].weight.detach().cpu(), w2[n], 2e-6, f"{n} w2 step {step}")
            assert_close_rel(emb.first_order_embeddings[n].weight.detach().cpu(), w1[n], 2e-6, f"{n} w1 step {step}")
    # moments of untouched rows stay zero, weights of untouched rows stay as initialised
    slots = dict(zip(emb._slot_of_param, range(len(emb._slot_of_param))))
    f_big = list(schema.fields).index("s1")
    ea = opt.exp_avg[slots[5 * f_big]].cpu().numpy()
    assert (np.abs(ea).sum(axis=1) == 0).sum() > 0 and np.array_equal(ea != 0, m2["s1"] != 0)


def test_first_step_equals_torch_adam_on_touched_rows():
    """One step from zero moments: torch.optim.Adam on the materialised sparse-as-dense gradient gives the same
    weights on the touched rows (untouched rows have zero gradient, so dense Adam leaves them alone too)."""
    torch.manual_seed(1)
    D, B = 16, 300
    schema = _schema(D)
    emb = FeatureEmbedding(schema, D).cuda()
    ref = FeatureEmbedding(schema, D).cuda()
    ref.load_state_dict(emb.state_dict())
    emb.grad_mode, ref.grad_mode = "row_sparse", "dense"
    batch = {"d0": torch.rand(B, device="cuda")}
    for n, f in schema.fields.items():
        if f.feature_type != FeatureType.DENSE:
            batch[n] = torch.randint(0, f.vocabulary_size, (B,), device="cuda")
    g = torch.randn(B, schema.total_embedding_dim, device="cuda")
    for m in (emb, ref):
        fo, fe, fl = m(batch)
        ((fl * g).sum() + fo.sum()).backward()
    RowSparseAdam(emb, lr=1e-3).step()
    tabs = [p for n, p in ref.named_parameters() if "embeddings.s" in n]
    torch.optim.Adam(tabs, lr=1e-3).step()
    for (n, p), (_, q) in zip(emb.named_parameters(), ref.named_parameters()):
        if "embeddings.s" in n:
            assert_close_rel(p.detach().cpu(), q.detach().cpu(), 2e-6, n)
