"""Drop-in ``RelGATLayer`` (same constructor, parameters, state-dict keys and forward signature
as reference relgat_projector/core/model/layer.py:9-323) running on the sm_100a kernels."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import functional as RF
from . import ops
from .graph import get_graph_index


class RelGATLayer(nn.Module):
    """Multi-head relational GAT layer.

    Parameter containers, shapes and the order in which the constructor consumes the RNG follow
    the reference (layer.py:108-129): ``proj`` = ModuleList of ``heads`` bias-free
    ``Linear(in_dim, out_dim)``, ``attn_vec`` = ParameterList of ``heads`` ``[num_rel, out_dim]``,
    ``rel_bias`` = ``[num_rel]`` zeros (or None); then xavier-uniform over ``proj`` and
    ``attn_vec``.  The same seed therefore yields the same weights and reference checkpoints
    load with ``strict=True``.

    ``precision``: "fp32" = tensor-core GEMMs on bf16 (hi, lo) splits reproducing fp32 to ~1e-5
    (parity mode); "bf16" = single-pass bf16 operands with fp32 accumulation and fp32 storage.
    """

    STABLE_SOFTMAX_EPS = 1e-16

    def __init__(
        self,
        in_dim: int,
        out_dim: int,
        num_rel: int,
        heads: int = 4,
        dropout: float = 0.2,
        use_bias: bool = True,
        relation_attn_dropout: Optional[float] = None,
        precision: str = "fp32",
    ):
        super().__init__()
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.heads = heads
        self.num_rel = num_rel
        self.precision = precision
        self.dropout = nn.Dropout(dropout)
        self.rel_attn_drop = nn.Dropout(p=relation_attn_dropout if relation_attn_dropout is not None else 0.0)
        self.proj = nn.ModuleList([nn.Linear(in_dim, out_dim, bias=False) for _ in range(heads)])
        self.attn_vec = nn.ParameterList([nn.Parameter(torch.empty(num_rel, out_dim)) for _ in range(heads)])
        self.rel_bias = nn.Parameter(torch.zeros(num_rel)) if use_bias else None
        self.reset_parameters()

    def reset_parameters(self):
        for lin in self.proj:
            nn.init.xavier_uniform_(lin.weight)
        for a in self.attn_vec:
            nn.init.xavier_uniform_(a)

    # -- packed views of the parameters the kernels consume -----------------------------------
    def packed_weight(self) -> torch.Tensor:
        """[heads*out_dim, in_dim]: one GEMM for all heads (column block h = proj[h])."""
        return torch.cat([lin.weight for lin in self.proj], dim=0)

    def packed_attention(self) -> torch.Tensor:
        """[heads, num_rel, out_dim]."""
        return torch.stack(list(self.attn_vec), dim=0)

    def kernel_params(self):
        return self.packed_weight(), self.packed_attention(), self.rel_bias

    def check_supported(self):
        """Kept for callers of round 1: every dropout configuration of the reference is supported now."""

    def draw_dropout(self, num_nodes: int, num_edges: int, device) -> Optional[RF.LayerDropout]:
        """Masks of this layer's two dropout sites for one step (None in eval mode or when both rates are 0):
        feature dropout on the output rows (reference layer.py:321-322) and attention dropout on alpha
        (layer.py:296-297).  Both are applied INSIDE the fused edge kernels; the bits come from Philox keyed by a
        seed drawn from torch's generator (``torch.manual_seed`` governs them)."""
        if not self.training:
            return None
        p_feat, p_attn = float(self.dropout.p), float(self.rel_attn_drop.p)
        if p_feat <= 0.0 and p_attn <= 0.0:
            return None
        C = self.heads * self.out_dim
        feat = ops.DropMask.draw((num_nodes, (C + 31) // 32), p_feat, device) if p_feat > 0.0 else None
        edge = ops.DropMask.draw(((num_edges * self.heads + 31) // 32,), p_attn, device) if p_attn > 0.0 else None
        return RF.LayerDropout(feat, edge)

    def forward(self, node_emb: torch.Tensor, edge_index: torch.Tensor, edge_type: torch.Tensor) -> torch.Tensor:
        """node_emb [N, in_dim] fp32, edge_index [2, E] int64 (row 0 = src, row 1 = dst),
        edge_type [E] int64  ->  [N, heads*out_dim] (column = h*out_dim + f)."""
        graph = get_graph_index(edge_index, edge_type, node_emb.size(0), self.num_rel)
        drop = self.draw_dropout(graph.N, graph.E, node_emb.device)
        return RF.relgat_stack(node_emb, graph, self.heads, self.out_dim, [self.kernel_params()],
                               precision=self.precision, drop=None if drop is None else [drop])
