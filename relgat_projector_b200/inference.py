"""Inference helpers for the inductive-imputation / query-expansion use of the model (BASELINE.json
config 5).  The reference only offers the hooks ``RelGATModel.get_node_repr`` and
``transform_from_vectors`` (reference core/model/model.py:144-186); its README describes masking
nodes and composing relation operators along a path without shipping code for either
(SURVEY.md §3.4).  These helpers are thin loops over those two hooks, on the GPU kernels."""
from __future__ import annotations

from typing import Sequence

import torch


@torch.no_grad()
def impute_masked_nodes(model, masked_ids: torch.Tensor) -> torch.Tensor:
    """Node representations when the input rows of ``masked_ids`` are unknown (zeroed): a masked node is
    rebuilt purely from its in-neighbours by the full-graph forward (+ projection head).  Returns the rows of
    the masked nodes, ``[len(masked_ids), D_sc]``.  The model's buffer is restored afterwards."""
    x = model.node_emb_fixed
    saved = x[masked_ids].clone()
    was_training = model.training
    model.eval()
    try:
        x[masked_ids] = 0  # bumps the version counter: the cached bf16 split of the inputs is rebuilt
        rows = model.get_node_repr()[masked_ids].clone()
    finally:
        x[masked_ids] = saved
        model.train(was_training)
    return rows


@torch.no_grad()
def expand_relation_path(model, src_vectors: torch.Tensor, rel_path: Sequence[int]) -> torch.Tensor:
    """Compose the scorer's relation operators along ``rel_path`` (query expansion): applies
    ``transform_from_vectors`` once per hop.  ``src_vectors`` lives in the scorer's space ``[B, D_sc]``."""
    v = src_vectors
    for r in rel_path:
        v = model.transform_from_vectors(v, torch.tensor([int(r)], device=v.device))
    return v
