"""``ProjectionHead`` with the reference's structure and state-dict keys (reference
relgat_projector/core/model/projection.py:7-72).  A plain dense MLP: left to cuBLAS/ATen
(SURVEY.md §2 row 4 — out of scope for hand kernels; "next" row §8(f)-1)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


class ProjectionHead(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, num_layers: int = 1, dropout: float = 0.0,
                 hidden_dim: Optional[int] = None):
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.hidden_dim = hidden_dim if hidden_dim is not None and hidden_dim > 0 else in_dim
        self.num_layers = max(0, int(num_layers))
        self.dropout = nn.Dropout(dropout) if dropout and dropout > 0 else nn.Identity()
        if self.num_layers == 0 and in_dim == out_dim:
            self.net = nn.Identity()
        elif self.num_layers <= 1:
            self.net = nn.Linear(in_dim, out_dim, bias=False)
        else:
            blocks = []
            width = in_dim
            for _ in range(self.num_layers - 1):  # Linear -> GELU -> LayerNorm blocks (projection.py:54-66)
                blocks += [nn.Linear(width, self.hidden_dim, bias=False), nn.GELU(), nn.LayerNorm(self.hidden_dim)]
                width = self.hidden_dim
            blocks.append(nn.Linear(self.hidden_dim, out_dim, bias=False))
            self.net = nn.Sequential(*blocks)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dropout(self.net(x))
