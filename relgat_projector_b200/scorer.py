"""Drop-in ``DistMultScorer`` / ``TransEScorer`` (reference relgat_projector/core/scorer.py:5-201)
on the fused gather-score kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as RF


class _ScorerBase(nn.Module):
    kind = ""

    def __init__(self, num_rel: int, rel_dim: int):
        super().__init__()
        self.rel_emb = nn.Embedding(num_rel, rel_dim)  # default normal init first (RNG order of the
        nn.init.xavier_uniform_(self.rel_emb.weight)   # reference, scorer.py:55-56), then xavier

    @property
    def _normalize(self) -> bool:
        return bool(getattr(self, "normalize", False))

    def forward(self, src_emb: torch.Tensor, rel_ids: torch.Tensor, dst_emb: torch.Tensor) -> torch.Tensor:
        """src_emb, dst_emb [B, D]; rel_ids [B] int64 -> scores [B] (higher = more plausible)."""
        score, _ = RF.ScoreRowsFunction.apply(self.kind, self._normalize, src_emb, dst_emb, self.rel_emb.weight,
                                              rel_ids, True, False)
        return score

    def transform(self, src_emb: torch.Tensor, rel_ids: torch.Tensor) -> torch.Tensor:
        """Relation operator applied to the source rows -> [B, D]."""
        _, tr = RF.ScoreRowsFunction.apply(self.kind, self._normalize, src_emb, None, self.rel_emb.weight,
                                           rel_ids, False, True)
        return tr

    def score_and_transform(self, src_emb: torch.Tensor, rel_ids: torch.Tensor, dst_emb: torch.Tensor,
                            n_transform: int = 0):
        """forward() and transform() of the first ``n_transform`` triples in ONE kernel launch."""
        return RF.ScoreRowsFunction.apply(self.kind, self._normalize, src_emb, dst_emb, self.rel_emb.weight, rel_ids,
                                          True, int(n_transform))

    def gather_score(self, x: torch.Tensor, src_ids, rel_ids, dst_ids, n_transform: int = 0,
                     want_dst_vec: bool = False):
        """Fused x[src_ids], x[dst_ids] gather + score (+ transform of the first n_transform triples,
        + gathered destination rows) — the model-level seam of reference model.py:135-141."""
        return RF.GatherScoreFunction.apply(self.kind, self._normalize, x, src_ids, dst_ids, self.rel_emb.weight,
                                            rel_ids, n_transform, want_dst_vec)


class DistMultScorer(_ScorerBase):
    """score = sum_d s*r*t ; transform = s (.) r   (reference scorer.py:80-84, 93-94)."""

    kind = "distmult"

    def __init__(self, num_rel: int, rel_dim: int):
        super().__init__(num_rel, rel_dim)


class TransEScorer(_ScorerBase):
    """score = -||s + r - t||_2, optionally on L2-normalised vectors (reference scorer.py:176-201)."""

    kind = "transe"

    def __init__(self, num_rel: int, rel_dim: int, normalize: bool = False):
        super().__init__(num_rel, rel_dim)
        self.normalize = normalize
