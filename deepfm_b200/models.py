"""CTR models wired exactly like the reference, on top of the CUDA drop-in layers.

Reference: ``BaseCTRModel`` deepfm/models/base.py:15-83; ``DeepFM`` deepfm.py:13-42; ``xDeepFM``
xdeepfm.py:13-48; ``AttentionDeepFM`` attention_deepfm.py:14-66; registry models/__init__.py:12-36.
Same constructor ``(schema, config)``, attribute names (hence ``state_dict`` keys), abstract hooks
``_build_components`` / ``_forward_components`` and ``forward`` / ``predict`` /
``get_l2_reg_loss``.  Only the layer classes differ: they run the sm_100a kernels.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, Type

import torch
import torch.nn as nn

from .layers.dnn import DNN, linear_head
from .layers.embedding import FeatureEmbedding
from .layers.fm import FMInteraction
from .layers.l2 import l2_penalty, prefetch_l2


class BaseCTRModel(nn.Module, ABC):
    # Optional hook (not in the reference): a callable (schema, fm_embed_dim) -> module used instead of
    # FeatureEmbedding, e.g. a row-sharded ShardedFeatureEmbedding for multi-GPU runs.
    embedding_factory = None

    def __init__(self, schema, config) -> None:
        super().__init__()
        self.schema = schema
        self.config = config
        factory = type(self).embedding_factory or FeatureEmbedding
        self.embedding = factory(schema, fm_embed_dim=config.feature.fm_embed_dim)
        self._build_components()

    @abstractmethod
    def _build_components(self) -> None:
        """Create the model-specific layers."""

    @abstractmethod
    def _forward_components(self, first_order: torch.Tensor, field_embeddings: torch.Tensor,
                            flat_embeddings: torch.Tensor) -> torch.Tensor:
        """Combine the three embedding views into raw logits (B, 1)."""

    def forward(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        first_order, field_embeddings, flat_embeddings = self.embedding(batch)
        if self.training and torch.is_grad_enabled():
            # the L2 value of get_l2_reg_loss() starts now, on a side stream, underneath the interaction / DNN
            # forward (after the HBM-bound embedding kernel, which it would only slow down)
            prefetch_l2(self.embedding, float(self.config.feature.embedding_l2_reg))
        return self._forward_components(first_order, field_embeddings, flat_embeddings)

    def predict(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        return torch.sigmoid(self.forward(batch))

    def get_l2_reg_loss(self) -> torch.Tensor:
        """``embedding_l2_reg * sum_p ||p||^2`` over every embedding parameter (base.py:78-83)."""
        return l2_penalty(self.embedding, self.config.feature.embedding_l2_reg)

    def _dnn(self, input_dim: int) -> DNN:
        c = self.config.dnn
        return DNN(input_dim=input_dim, hidden_units=c.hidden_units, activation=c.activation,
                   dropout=c.dropout, use_batch_norm=c.use_batch_norm)


class DeepFM(BaseCTRModel):
    """logit = first_order + FM(field_embeddings) + Linear(DNN(flat_embeddings))."""

    def _build_components(self) -> None:
        self.fm = FMInteraction()
        self.dnn = self._dnn(self.schema.total_embedding_dim)
        self.output_linear = nn.Linear(self.dnn.output_dim, 1)

    def _forward_components(self, first_order, field_embeddings, flat_embeddings):
        return first_order + self.fm(field_embeddings) + linear_head(self.output_linear, self.dnn(flat_embeddings))


class xDeepFM(BaseCTRModel):
    """logit = first_order + Linear(CIN(field_embeddings)) + Linear(DNN(flat_embeddings))."""

    def _build_components(self) -> None:
        from .layers.cin import CIN
        self.cin = CIN(num_fields=self.schema.num_fields, embed_dim=self.config.feature.fm_embed_dim,
                       layer_sizes=self.config.cin.layer_sizes, split_half=self.config.cin.split_half)
        self.dnn = self._dnn(self.schema.total_embedding_dim)
        self.cin_linear = nn.Linear(self.cin.output_dim, 1)
        self.dnn_linear = nn.Linear(self.dnn.output_dim, 1)

    def _forward_components(self, first_order, field_embeddings, flat_embeddings):
        return (first_order + linear_head(self.cin_linear, self.cin(field_embeddings))
                + linear_head(self.dnn_linear, self.dnn(flat_embeddings)))


class AttentionDeepFM(BaseCTRModel):
    """logit = first_order + FM(e) + Linear(DNN(cat(Attention(e).flatten(), flat)))."""

    def _build_components(self) -> None:
        from .layers.attention import MultiHeadSelfAttention
        a = self.config.attention
        self.fm = FMInteraction()
        self.attention = MultiHeadSelfAttention(embed_dim=self.config.feature.fm_embed_dim,
                                                num_heads=a.num_heads, attention_dim=a.attention_dim,
                                                num_layers=a.num_layers, use_residual=a.use_residual)
        width = self.schema.num_fields * self.config.feature.fm_embed_dim + self.schema.total_embedding_dim
        self.dnn = self._dnn(width)
        self.output_linear = nn.Linear(self.dnn.output_dim, 1)

    def _forward_components(self, first_order, field_embeddings, flat_embeddings):
        fm_out = self.fm(field_embeddings)
        refined = self.attention(field_embeddings)
        tower_in = torch.cat([refined.reshape(refined.size(0), -1), flat_embeddings], dim=1)
        return first_order + fm_out + linear_head(self.output_linear, self.dnn(tower_in))


MODEL_REGISTRY: Dict[str, Type[BaseCTRModel]] = {
    "deepfm": DeepFM,
    "xdeepfm": xDeepFM,
    "attention_deepfm": AttentionDeepFM,
}


def create_model(name: str, schema, config) -> BaseCTRModel:
    if name not in MODEL_REGISTRY:
        raise ValueError(f"Unknown model: {name}. Choose from {list(MODEL_REGISTRY)}")
    return MODEL_REGISTRY[name](schema, config)
