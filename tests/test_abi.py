"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol declared in
include/deepfm_b200.h, the host-only plan functions agree with the oracle's integer layout, the
drop-in modules keep the reference's state_dict keys, and nothing silently falls back to CPU."""

import ctypes as C

import numpy as np
import pytest
import torch

from deepfm_b200 import _lib
from oracle import deepfm_oracle as O
from tests.conftest import load_golden, split_prefixed
from tests.golden import spec


def test_library_exports_every_header_symbol():
    names = _lib.header_symbols()
    assert len(names) >= 16
    handle = _lib.lib()
    missing = [n for n in names if not hasattr(handle, n)]
    assert not missing, missing
    assert set(names) <= set(_lib.SIGNATURES), sorted(set(names) - set(_lib.SIGNATURES))
    assert handle.dfm_version() >= 100


def _plan(schema, D):
    from deepfm_b200.layers.embedding import FeatureEmbedding
    emb = FeatureEmbedding(schema, fm_embed_dim=D)
    return emb, emb._ensure_plan()


def test_plan_sizes_and_slot_layout_match_oracle():
    schema = spec.golden_schema()
    emb, plan = _plan(schema, spec.FM_DIM)
    info = (C.c_int64 * 8)()
    assert _lib.lib().dfm_plan_info(plan, info) == 0
    sf, sp, rb = O.slot_layout(schema)
    assert info[0] == schema.total_embedding_dim and info[1] == len(sf) and info[2] == rb[-1]
    assert info[4] == 0 and info[5] == 8 and info[7] == 4
    assert (1 << info[6]) > rb[-1]                      # PAD key (= total rows) is sortable
    a = (C.c_int32 * len(sf))()
    b = (C.c_int32 * len(sf))()
    c = (C.c_int64 * (len(schema.fields) + 1))()
    assert _lib.lib().dfm_plan_slots(plan, a, b, c) == 0
    assert list(a) == sf.tolist() and list(b) == sp.tolist() and list(c) == rb.tolist()


def test_plan_create_rejects_bad_arguments():
    lib = _lib.lib()
    bad = lib.dfm_plan_create(1, _lib.i32_array([7]), _lib.i32_array([4]), _lib.i64_array([5]),
                              _lib.i32_array([1]), _lib.i32_array([0]), 4)
    assert not bad and "invalid" in _lib.last_error()
    assert not lib.dfm_plan_create(0, None, None, None, None, None, 4)


@pytest.mark.parametrize("name", ["deepfm", "xdeepfm", "attention_deepfm"])
def test_models_keep_reference_state_dict_keys_and_shapes(name):
    from deepfm_b200.models import create_model
    g = load_golden(f"model_{name}.npz")
    ref = split_prefixed(g, "param/")
    model = create_model(name, spec.golden_schema(), spec.golden_config())
    sd = model.state_dict()
    assert set(sd) == set(ref)
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k].shape), k
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in ref.items()})
    emb_keys = [k for k, _ in model.embedding.named_parameters()]
    assert emb_keys == [k[len("embedding."):] for k in ref if k.startswith("embedding.")
                        and "num_batches" not in k][:len(emb_keys)] or set(emb_keys) == \
        {k[len("embedding."):] for k in ref if k.startswith("embedding.")}


def test_unknown_model_raises_value_error():
    from deepfm_b200.models import create_model
    with pytest.raises(ValueError):
        create_model("nope", spec.golden_schema(), spec.golden_config())


def test_attention_dim_must_divide_heads():
    from deepfm_b200.layers.attention import MultiHeadSelfAttention
    with pytest.raises(ValueError):
        MultiHeadSelfAttention(embed_dim=8, num_heads=3, attention_dim=8)


def test_embedding_init_matches_reference_convention():
    from deepfm_b200.layers.embedding import FeatureEmbedding
    emb = FeatureEmbedding(spec.golden_schema(), fm_embed_dim=spec.FM_DIM)
    for name, f in emb.schema.fields.items():
        w = emb.second_order_embeddings[name].weight
        if O._kind(f) != "dense":
            assert torch.all(w[0] == 0) and torch.any(w[1:] != 0)
        else:
            assert torch.all(emb.second_order_embeddings[name].bias == 0)
    assert set(emb.projections.keys()) == {"i", "g", "m", "y"}


def test_cpu_tensors_are_refused_not_silently_computed():
    from deepfm_b200.layers.embedding import FeatureEmbedding
    from deepfm_b200.layers.fm import FMInteraction
    emb = FeatureEmbedding(spec.golden_schema(), fm_embed_dim=spec.FM_DIM)
    batch = {k: torch.from_numpy(v) for k, v in spec.golden_batch().items()}
    with pytest.raises(RuntimeError, match="no CPU"):
        emb(batch)
    with pytest.raises(RuntimeError, match="no CPU"):
        FMInteraction()(torch.randn(2, 3, 4))


def test_tower_store_layout_and_workspaces_are_consistent():
    """Host-only entry points of the sequenced DNN tower (no GPU work): the activation store holds, per block, the
    pre-activation, the block output (not for the last block: that is the caller's `out`) and the batch statistics, every
    piece 256-byte aligned and non-overlapping; workspaces cover every product of every block."""
    lib = _lib.lib()
    dims = [2496, 256, 128, 64]
    M, n = 4096, 3
    c_dims = _lib.i64_array(dims)
    offs = (C.c_int64 * (4 * n))()
    total = lib.dfm_tower_store_floats(n, c_dims, M, offs)
    assert total == lib.dfm_tower_store_floats(n, c_dims, M, None)
    pieces = []
    for l in range(n):
        h = dims[l + 1]
        oy, oa, om, orr = (offs[4 * l + k] for k in range(4))
        assert (oa == -1) == (l == n - 1)
        pieces += [(oy, M * h), (om, h), (orr, h)] + ([(oa, M * h)] if oa >= 0 else [])
    pieces.sort()
    for (o0, n0), (o1, _) in zip(pieces, pieces[1:]):
        assert o0 % 64 == 0 and o0 + n0 <= o1
    assert pieces[-1][0] + pieces[-1][1] <= total
    fwd = lib.dfm_tower_seq_workspace_bytes(n, c_dims, M, 0)
    bwd = lib.dfm_tower_seq_workspace_bytes(n, c_dims, M, 1)
    need = max(lib.dfm_gemm3_workspace_bytes(0, M, dims[l + 1], dims[l]) + lib.dfm_tower_workspace_bytes(M, dims[l + 1]) for l in range(n))
    assert fwd >= need
    need_b = max(max(lib.dfm_gemm3_workspace_bytes(1, M, dims[l], dims[l + 1]), lib.dfm_gemm3_workspace_bytes(2, dims[l + 1], dims[l], M))
                 for l in range(n)) + 3 * M * max(dims[1:]) * 4
    assert bwd >= need_b and bwd > fwd
    # mode 2 (weight gradient) runs swapped: its workspace holds the split-K partials AND the lo part of dY (K x M floats)
    assert lib.dfm_gemm3_workspace_bytes(2, 256, 2496, 65536) >= 65536 * 256 * 4
