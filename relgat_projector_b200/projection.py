"""``ProjectionHead`` with the reference's structure and state-dict keys (reference
relgat_projector/core/model/projection.py:7-72).  "Next" row §8(f)-1: on CUDA its bias-free linears
run on the tcgen05 GEMM of the hot path (fp32-accurate bf16 hi/lo split, or single-pass bf16),
because torch's fp32 SGEMM over all N rows would cost more than the whole GAT stack; GELU and
LayerNorm stay ATen element-wise ops."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


class ProjectionHead(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, num_layers: int = 1, dropout: float = 0.0,
                 hidden_dim: Optional[int] = None, precision: str = "fp32"):
        super().__init__()
        self.precision = precision
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

    def _run(self, mod: nn.Module, x: torch.Tensor) -> torch.Tensor:
        if isinstance(mod, nn.Linear) and mod.bias is None and x.is_cuda and x.dtype == torch.float32:
            from .functional import split_linear
            return split_linear(x, mod.weight, self.precision)
        return mod(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if isinstance(self.net, nn.Sequential):
            for mod in self.net:
                x = self._run(mod, x)
        else:
            x = self._run(self.net, x)
        return self.dropout(x)
