"""Shared description of the golden cases (schema, configs, batches).

Used by ``make_golden.py`` (build container only: imports /root/reference) and by the tests
(everywhere: reads the committed ``*.npz``).  Nothing here touches the reference.
"""

from __future__ import annotations

import numpy as np

from deepfm_b200.config import ExperimentConfig
from deepfm_b200.schema import DatasetSchema, FeatureType, FieldSchema

FM_DIM = 8
L2_REG = 1e-3  # larger than the 1e-5 default so the L2 term is visible in the gradients


def golden_schema(cls_schema=DatasetSchema, cls_field=FieldSchema, ftype=FeatureType):
    """Every field kind, every combiner, mixed dims (some projected to FM_DIM, some not)."""
    S, Q, N = ftype.SPARSE, ftype.SEQUENCE, ftype.DENSE
    fields = {
        "u": cls_field("u", S, vocabulary_size=11, embedding_dim=8),
        "i": cls_field("i", S, vocabulary_size=7, embedding_dim=4),
        "g": cls_field("g", Q, vocabulary_size=9, embedding_dim=4, max_length=5, combiner="mean"),
        "x": cls_field("x", N, embedding_dim=8),
        "t": cls_field("t", Q, vocabulary_size=6, embedding_dim=8, max_length=3, combiner="sum"),
        "m": cls_field("m", Q, vocabulary_size=6, embedding_dim=4, max_length=4, combiner="max"),
        "y": cls_field("y", N, embedding_dim=4),
    }
    return cls_schema(fields=fields, label_field="label")


def golden_batch():
    """B=6; includes id 0 in a sparse field, mid-sequence pads, duplicates, an all-pad bag."""
    return {
        "u": np.array([3, 0, 10, 3, 1, 7], dtype=np.int64),
        "i": np.array([1, 6, 6, 2, 0, 5], dtype=np.int64),
        "g": np.array([[1, 2, 0, 0, 0], [0, 0, 0, 0, 0], [8, 0, 8, 3, 0],
                       [4, 4, 4, 4, 4], [0, 0, 0, 0, 7], [2, 0, 5, 0, 1]], dtype=np.int64),
        "x": np.array([0.5, -1.0, 0.25, 0.0, 0.75, -0.125], dtype=np.float32),
        "t": np.array([[1, 1, 0], [5, 0, 2], [0, 0, 0], [3, 4, 5], [0, 2, 0], [2, 2, 2]], dtype=np.int64),
        "m": np.array([[1, 2, 3, 0], [0, 0, 0, 0], [5, 5, 0, 1], [0, 4, 0, 0],
                       [2, 3, 4, 5], [1, 0, 0, 1]], dtype=np.int64),
        "y": np.array([1.0, 0.5, -0.5, 0.125, 0.0, 2.0], dtype=np.float32),
    }


def golden_labels():
    return np.array([1, 0, 0, 1, 0, 1], dtype=np.float32)


def golden_config(cls=ExperimentConfig):
    cfg = cls()
    cfg.feature.fm_embed_dim = FM_DIM
    cfg.feature.embedding_l2_reg = L2_REG
    cfg.dnn.hidden_units = [16, 8]
    cfg.dnn.dropout = 0.0
    cfg.cin.layer_sizes = [6, 5, 4]      # odd size: direct 2 / next 3
    cfg.cin.split_half = True
    cfg.attention.num_heads = 2
    cfg.attention.attention_dim = 8
    cfg.attention.num_layers = 2
    cfg.attention.use_residual = True
    return cfg
