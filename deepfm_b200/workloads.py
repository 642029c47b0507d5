"""Synthetic workloads of BASELINE.json's configs (shapes only; there is no dataset in the image).

``ml100k_schema``  -- the 16-field MovieLens-100K schema the reference's adapter builds
                      (deepfm/data/movielens.py:346-418; vocabulary sizes from the reference's
                      notebooks/feature_embedding_guide.ipynb cell 4, bucket fields estimated).
``criteo_schema``  -- 13 DENSE + 26 SPARSE fields, embedding_dim == fm_embed_dim (no projections),
                      Criteo-Kaggle cardinalities (sum = 33.76 M rows).
``criteo_multihot_schema`` -- the 26 categorical fields as SEQUENCE bags (config 5).
Id columns are log-uniform over [1, V) (P(id = k) ~ 1/k, i.e. Zipf with alpha = 1): a few hot
rows per table plus a long tail, like CTR logs.
"""

from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from .schema import DatasetSchema, FeatureType, FieldSchema

CRITEO_VOCAB = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27,
                14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]


def ml100k_schema() -> DatasetSchema:
    S, Q, N = FeatureType.SPARSE, FeatureType.SEQUENCE, FeatureType.DENSE
    spec = [("user_id", S, 944, 16), ("movie_id", S, 1679, 16), ("gender", S, 3, 4), ("age", S, 8, 4),
            ("occupation", S, 22, 8), ("zip_prefix", S, 383, 8), ("genres", Q, 20, 8),
            ("release_year_bucket", S, 18, 4), ("movie_age_at_rating", S, 8, 4), ("num_genres", S, 7, 4),
            ("dow_sin", N, 0, 4), ("dow_cos", N, 0, 4), ("hour_sin", N, 0, 4), ("hour_cos", N, 0, 4),
            ("user_rating_count", N, 0, 8), ("item_rating_count", N, 0, 8)]
    fields = {}
    for name, kind, vocab, dim in spec:
        fields[name] = FieldSchema(name, kind, vocabulary_size=vocab, embedding_dim=dim,
                                   max_length=6 if kind == Q else 1, combiner="mean")
    return DatasetSchema(fields=fields, label_field="label")


def criteo_schema(embed_dim: int = 64, vocab_scale: float = 1.0, max_vocab: Optional[int] = None) -> DatasetSchema:
    fields = {}
    for i in range(13):
        fields[f"I{i + 1}"] = FieldSchema(f"I{i + 1}", FeatureType.DENSE, embedding_dim=embed_dim)
    for i, v in enumerate(CRITEO_VOCAB):
        v = max(3, int(v * vocab_scale))
        if max_vocab:
            v = min(v, max_vocab)
        fields[f"C{i + 1}"] = FieldSchema(f"C{i + 1}", FeatureType.SPARSE, vocabulary_size=v, embedding_dim=embed_dim)
    return DatasetSchema(fields=fields, label_field="label")


def criteo_multihot_schema(embed_dim: int = 64, max_length: int = 16, vocab_scale: float = 1.0,
                           combiner: str = "mean") -> DatasetSchema:
    fields = {}
    for i in range(13):
        fields[f"I{i + 1}"] = FieldSchema(f"I{i + 1}", FeatureType.DENSE, embedding_dim=embed_dim)
    for i, v in enumerate(CRITEO_VOCAB):
        fields[f"C{i + 1}"] = FieldSchema(f"C{i + 1}", FeatureType.SEQUENCE, vocabulary_size=max(3, int(v * vocab_scale)),
                                          embedding_dim=embed_dim, max_length=max_length, combiner=combiner)
    return DatasetSchema(fields=fields, label_field="label")


def synthetic_batch(schema: DatasetSchema, batch: int, seed: int = 0, device="cpu", avg_nnz: float = 8.0
                    ) -> Dict[str, torch.Tensor]:
    """Columnar batch with the dtypes/layouts the reference's TabularDataset yields
    (deepfm/data/dataset.py:28-38): int64 (B,) / (B, L) zero-padded ids, float32 (B,) dense."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for name, f in schema.fields.items():
        kind = f.feature_type.value if hasattr(f.feature_type, "value") else str(f.feature_type)
        if kind == "dense":
            out[name] = torch.rand(batch, generator=gen) * 2 - 1
            continue
        V = f.vocabulary_size
        shape = (batch,) if kind == "sparse" else (batch, f.max_length)
        u = torch.rand(shape, generator=gen, dtype=torch.float64)
        ids = torch.exp(u * math.log(max(V - 1, 1))).floor().long().clamp_(1, max(V - 1, 1))
        if kind == "sequence":
            L = f.max_length
            nnz = torch.poisson(torch.full((batch,), float(min(avg_nnz, L)) - 1.0), generator=gen).long().clamp_(0, L - 1) + 1
            ids = ids * (torch.arange(L)[None, :] < nnz[:, None])
        out[name] = ids
    return {k: v.to(device) for k, v in out.items()}


def synthetic_labels(batch: int, seed: int = 0, device="cpu", positive_rate: float = 0.2) -> torch.Tensor:
    gen = torch.Generator(device="cpu").manual_seed(seed + 1)
    return (torch.rand(batch, generator=gen) < positive_rate).float().to(device)
