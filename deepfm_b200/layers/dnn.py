"""MLP tower (reference: deepfm/models/layers/dnn.py:8-59).  Out of custom-kernel scope
(SURVEY section 2): it stays ``nn.Linear`` -> cuBLAS and is the data-parallel allreduce payload.
Same constructor, ``mlp`` Sequential layout (hence ``state_dict`` keys) and ``output_dim``."""

from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

_ACTIVATIONS = {"relu": nn.ReLU, "leaky_relu": nn.LeakyReLU, "gelu": nn.GELU, "tanh": nn.Tanh}


class DNN(nn.Module):
    ACTIVATIONS = _ACTIVATIONS

    def __init__(self, input_dim: int, hidden_units: List[int], activation: str = "relu",
                 dropout: float = 0.1, use_batch_norm: bool = True) -> None:
        super().__init__()
        if not hidden_units:
            raise ValueError("hidden_units must be non-empty")
        act = _ACTIVATIONS.get(activation.lower())
        if act is None:
            raise ValueError(f"Unknown activation: {activation}. Choose from {list(_ACTIVATIONS)}")
        blocks: List[nn.Module] = []
        width = input_dim
        for units in hidden_units:
            blocks.append(nn.Linear(width, units))
            if use_batch_norm:
                blocks.append(nn.BatchNorm1d(units))
            blocks += [act(), nn.Dropout(p=dropout)]
            width = units
        self.mlp = nn.Sequential(*blocks)
        self.output_dim = hidden_units[-1]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.mlp(x)
