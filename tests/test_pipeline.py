"""Device-side input pipeline (deepfm_b200/pipeline.py) against the reference's dataset conventions
(deepfm/data/dataset.py:28-38, tests/test_dataset.py:37-73: int -> long, float -> float32, 2-D sequence features) and
the oracle's CSR flatten (bit-exact).  CPU part runs anywhere; the GPU part checks the prefetching loader."""

import numpy as np
import pytest
import torch

from deepfm_b200.pipeline import ColumnarDataset, DeviceLoader, to_csr
from oracle import deepfm_oracle as O


def _data(n=1000, seed=0):
    rng = np.random.default_rng(seed)
    feats = {"u": rng.integers(0, 50, n).astype(np.int32), "x": rng.random(n).astype(np.float64),
             "g": (rng.integers(0, 9, (n, 5)) * (rng.random((n, 5)) < 0.6)).astype(np.int64)}
    return feats, (rng.random(n) < 0.3).astype(np.int64)


def test_columnar_dataset_dtypes_and_batches_match_reference_conventions():
    feats, labels = _data()
    ds = ColumnarDataset(feats, labels, pin=False)
    assert len(ds) == 1000
    loader = DeviceLoader(ds, batch_size=256, shuffle=False)
    assert len(loader) == 4                                        # drop_last=False like the reference DataLoader
    seen = 0
    for bf, bl in loader:
        assert bf["u"].dtype == torch.long and bf["g"].dtype == torch.long and bf["x"].dtype == torch.float32
        assert bl.dtype == torch.float32 and bf["g"].dim() == 2 and bf["g"].shape[1] == 5
        b = bl.shape[0]
        assert np.array_equal(bf["u"].numpy(), feats["u"][seen:seen + b])
        assert np.array_equal(bf["g"].numpy(), feats["g"][seen:seen + b])
        seen += b
    assert seen == 1000
    row_f, row_l = ds[3]
    assert row_f["u"].item() == feats["u"][3] and row_l.item() == float(labels[3])


def test_shuffle_is_a_permutation_and_seeded():
    feats, labels = _data()
    ds = ColumnarDataset(feats, labels, pin=False)
    a = torch.cat([bf["u"] * 1000 + bf["g"][:, 0] for bf, _ in DeviceLoader(ds, 128, shuffle=True, seed=7)])
    b = torch.cat([bf["u"] * 1000 + bf["g"][:, 0] for bf, _ in DeviceLoader(ds, 128, shuffle=True, seed=7)])
    c = torch.cat([bf["u"] * 1000 + bf["g"][:, 0] for bf, _ in DeviceLoader(ds, 128, shuffle=True, seed=8)])
    assert torch.equal(a, b) and not torch.equal(a, c)
    want = torch.from_numpy(feats["u"].astype(np.int64) * 1000 + feats["g"][:, 0])
    assert torch.equal(torch.sort(a).values, torch.sort(want).values)


def test_to_csr_is_bit_exact_vs_oracle():
    feats, _ = _data(n=513, seed=3)
    vals, offs = to_csr(torch.from_numpy(feats["g"]))
    oo, ov = O.csr_flatten(feats["g"])          # oracle returns (offsets, values)
    assert np.array_equal(vals.numpy(), ov) and np.array_equal(offs.numpy(), oo)
    with pytest.raises(ValueError):
        to_csr(torch.zeros(4, dtype=torch.long))


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [True, False])
def test_loader_feeds_the_embedding_and_presorts_on_gpu(resident):
    """Host-resident (pinned, prefetched copies) and device-resident datasets give the same batches; with an embedding
    attached every batch arrives with its keys already sorted (the backward skips its sort) and gradients are unchanged."""
    from deepfm_b200.layers.embedding import FeatureEmbedding
    from deepfm_b200 import workloads as W
    schema = W.criteo_schema(64, vocab_scale=0.001)
    n, B = 6000, 2048
    host = W.synthetic_batch(schema, n, seed=1)
    ds = ColumnarDataset({k: v.numpy() for k, v in host.items()}, W.synthetic_labels(n, seed=1).numpy(),
                         device="cuda" if resident else None)
    torch.manual_seed(0)
    emb = FeatureEmbedding(schema, 64).cuda()
    emb.grad_mode = "row_sparse"
    loader = DeviceLoader(ds, B, shuffle=False, device="cuda", embedding=emb)
    lo = 0
    for bf, bl in loader:
        hi = min(lo + B, n)
        assert all(v.is_cuda for v in bf.values()) and bl.is_cuda
        assert torch.equal(bf["C3"].cpu(), host["C3"][lo:hi])
        fo, fe, fl = emb(bf)
        (fl.sum() + fo.sum()).backward()
        per_a = {k: tuple(t.clone() for t in v) for k, v in emb.row_grads.per_table().items()}
        emb.async_sort = False                       # same batch, sort inside the backward
        fo, fe, fl = emb({k: v.clone() for k, v in bf.items()})
        (fl.sum() + fo.sum()).backward()
        emb.async_sort = True
        for k, v in emb.row_grads.per_table().items():
            assert all(torch.equal(a, b) for a, b in zip(per_a[k], v)), k
        lo = hi
    assert lo == n


def test_packed_batch_layout_round_trips_every_column():
    """One contiguous buffer per batch (a single host -> device copy): the column views keep the reference's dtypes
    and shapes (dataset.py:28-38) and start on 256-byte boundaries."""
    from deepfm_b200.pipeline import PackedBatchLayout
    from deepfm_b200 import workloads as W
    schema = W.ml100k_schema()
    batch = W.synthetic_batch(schema, 37, seed=3)
    labels = W.synthetic_labels(37, seed=3)
    layout = PackedBatchLayout(batch, labels)
    buf = layout.pack(batch, labels, pin=False)
    assert buf.dtype == torch.uint8 and buf.numel() == layout.nbytes
    assert all(off % 256 == 0 for _, _, _, off, _ in layout.columns)
    feats, lab = layout.views(buf.clone())
    assert list(feats) == list(batch)
    for k, v in batch.items():
        assert feats[k].dtype == v.dtype and feats[k].shape == v.shape and torch.equal(feats[k], v)
    assert torch.equal(lab, labels)
    assert layout.payload_bytes == sum(v.numel() * v.element_size() for v in batch.values()) + labels.numel() * 4
