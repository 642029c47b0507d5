"""B200-native DeepFM embedding + interaction hot path (drop-in for CodexploreRepo/deepfm's
FeatureEmbedding / FMInteraction / CIN / MultiHeadSelfAttention and the BaseCTRModel wiring)."""

from .config import ExperimentConfig, load_config  # noqa: F401
from .schema import DatasetSchema, FeatureType, FieldSchema  # noqa: F401

__all__ = ["ExperimentConfig", "load_config", "DatasetSchema", "FeatureType", "FieldSchema"]
