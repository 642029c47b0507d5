"""Layer modules of the hot path (same names as deepfm/models/layers/__init__.py:3-15)."""

from .attention import MultiHeadSelfAttention
from .cin import CIN
from .dnn import DNN
from .embedding import FeatureEmbedding, RowSparseGrads
from .fm import FMInteraction

__all__ = ["CIN", "DNN", "FeatureEmbedding", "FMInteraction", "MultiHeadSelfAttention", "RowSparseGrads"]
