"""Experiment configuration read by the hot path.

Same nested dataclass tree, attribute names and defaults as the reference's
``deepfm/config.py:13-86`` so ``BaseCTRModel(schema, config)`` can be built from either
package's config object.  The reference needs the third-party ``dacite`` to hydrate YAML;
here ``from_dict`` is a 20-line recursive constructor, so there is no extra dependency.
Only ``feature.*``, ``cin.*``, ``attention.*``, ``dnn.*`` and ``training.*`` are read by
code in this repository (reference: base.py:32-34,83; xdeepfm.py:20-25;
attention_deepfm.py:27-33; trainer.py:59-78,232-237).
"""

from __future__ import annotations

import ast
import dataclasses
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, List, Optional


@dataclass
class DataConfig:
    dataset_name: str = "movielens"
    data_dir: str = "data/ml-100k"
    split_strategy: str = "temporal"
    temporal_val_ratio: float = 0.1
    temporal_test_ratio: float = 0.1
    neg_sampling_alpha: float = 0.75
    min_interactions: int = 3
    label_threshold: float = 4.0
    num_neg_train: int = 4
    num_neg_eval: int = 999


@dataclass
class FeatureConfig:
    fm_embed_dim: int = 16
    embedding_l2_reg: float = 1e-5


@dataclass
class FMConfig:
    use_first_order: bool = True
    use_second_order: bool = True


@dataclass
class DNNConfig:
    hidden_units: List[int] = field(default_factory=lambda: [256, 128, 64])
    activation: str = "relu"
    dropout: float = 0.1
    use_batch_norm: bool = True


@dataclass
class CINConfig:
    layer_sizes: List[int] = field(default_factory=lambda: [128, 128])
    split_half: bool = True


@dataclass
class AttentionConfig:
    num_heads: int = 4
    attention_dim: int = 64
    num_layers: int = 1
    use_residual: bool = True


@dataclass
class TrainingConfig:
    num_epochs: int = 50
    batch_size: int = 4096
    lr: float = 1e-3
    optimizer: str = "adam"
    scheduler: str = "reduce_on_plateau"
    early_stopping_patience: int = 5
    metric: str = "auc"
    gradient_clip_norm: float = 1.0
    ranking_ks: List[int] = field(default_factory=lambda: [1, 5, 10, 20])


@dataclass
class ExperimentConfig:
    model_name: str = "deepfm"
    seed: int = 42
    device: str = "auto"
    output_dir: str = "outputs"
    data: DataConfig = field(default_factory=DataConfig)
    feature: FeatureConfig = field(default_factory=FeatureConfig)
    fm: FMConfig = field(default_factory=FMConfig)
    dnn: DNNConfig = field(default_factory=DNNConfig)
    cin: CINConfig = field(default_factory=CINConfig)
    attention: AttentionConfig = field(default_factory=AttentionConfig)
    training: TrainingConfig = field(default_factory=TrainingConfig)


def from_dict(cls, data: dict):
    """Hydrate a (nested) dataclass from a plain dict; unknown keys raise ``ValueError``."""
    if not dataclasses.is_dataclass(cls):
        return data
    known = {f.name: f for f in dataclasses.fields(cls)}
    kwargs = {}
    for key, value in (data or {}).items():
        if key not in known:
            raise ValueError(f"unknown config key {key!r} for {cls.__name__}")
        ftype = known[key].type
        target = globals().get(ftype) if isinstance(ftype, str) else ftype
        if dataclasses.is_dataclass(target) and isinstance(value, dict):
            kwargs[key] = from_dict(target, value)
        else:
            kwargs[key] = value
    return cls(**kwargs)


def parse_override_value(text: str) -> Any:
    """``"true"``→bool, ints, floats, ``"[1,2]"``→list, else the string (config.py:113-131)."""
    low = text.lower()
    if low in ("true", "false"):
        return low == "true"
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            continue
    if text.startswith("[") and text.endswith("]"):
        try:
            return ast.literal_eval(text)
        except (ValueError, SyntaxError):
            pass
    return text


def load_config(yaml_path, overrides: Optional[List[str]] = None) -> ExperimentConfig:
    """YAML → ``ExperimentConfig`` with dotted ``a.b=c`` overrides (config.py:89-110)."""
    import yaml

    raw = yaml.safe_load(Path(yaml_path).read_text()) or {}
    for item in overrides or []:
        dotted, value = item.split("=", 1)
        node = raw
        *parents, leaf = dotted.strip().split(".")
        for p in parents:
            node = node.setdefault(p, {})
        node[leaf] = parse_override_value(value.strip())
    return from_dict(ExperimentConfig, raw)
