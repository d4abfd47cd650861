"""``ProjectionHead`` with the reference's structure and state-dict keys (reference
relgat_projector/core/model/projection.py:7-72).  "Next" row §8(f)-1: on CUDA its bias-free linears
run on the tcgen05 GEMM of the hot path (fp32-accurate bf16 hi/lo split, or single-pass bf16),
because torch's fp32 SGEMM over all N rows would cost more than the whole GAT stack; the GELU + LayerNorm
between them is one fused kernel per direction (csrc/gelu_ln.cu)."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn


class ProjectionHead(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, num_layers: int = 1, dropout: float = 0.0,
                 hidden_dim: Optional[int] = None, precision: str = "fp32"):
        super().__init__()
        self.precision = precision
        depth = max(0, int(num_layers))
        width = hidden_dim if (hidden_dim is not None and hidden_dim > 0) else in_dim
        self.in_dim, self.out_dim, self.hidden_dim, self.num_layers = in_dim, out_dim, width, depth
        self.dropout = nn.Dropout(dropout) if (dropout and dropout > 0) else nn.Identity()
        if depth == 0 and in_dim == out_dim:
            self.net = nn.Identity()
            return
        # chain of bias-free linears in_dim -> hidden ... -> out_dim; GELU + LayerNorm after every linear but the
        # last.  Module order (hence state-dict keys net.0, net.3, ... and the init-RNG order) as in the reference.
        dims = [in_dim] + [width] * max(depth - 1, 0) + [out_dim]
        mods = []
        for k in range(len(dims) - 1):
            mods.append(nn.Linear(dims[k], dims[k + 1], bias=False))
            if k < len(dims) - 2:
                mods += [nn.GELU(), nn.LayerNorm(dims[k + 1])]
        self.net = mods[0] if len(mods) == 1 else nn.Sequential(*mods)

    def _run(self, mod: nn.Module, x: torch.Tensor) -> torch.Tensor:
        if isinstance(mod, nn.Linear) and mod.bias is None and x.is_cuda and x.dtype == torch.float32:
            from .functional import split_linear
            return split_linear(x, mod.weight, self.precision)
        return mod(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if isinstance(self.net, nn.Sequential):
            mods = list(self.net)
            i = 0
            while i < len(mods):
                mod = mods[i]
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if (isinstance(mod, nn.GELU) and getattr(mod, "approximate", "none") == "none"
                        and isinstance(nxt, nn.LayerNorm) and len(nxt.normalized_shape) == 1
                        and x.is_cuda and x.dtype == torch.float32):
                    # GELU + LayerNorm of a hidden block: one fused kernel per direction (csrc/gelu_ln.cu)
                    from .functional import gelu_layernorm
                    x = gelu_layernorm(x, nxt.weight, nxt.bias, nxt.eps)
                    i += 2
                    continue
                x = self._run(mod, x)
                i += 1
        else:
            x = self._run(self.net, x)
        return self.dropout(x)
