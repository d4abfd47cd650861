"""Ranking and reconstruction losses over the scorer outputs (reference
relgat_projector/core/loss/{relgat_loss,multi_objective_loss,cosine,mse}.py).

Class names, constructor arguments and call signatures match the reference, so the reference trainer can use
either implementation.  CUDA float32 inputs run on the fused loss kernels of csrc/loss.cu (value and gradient in one
launch, fixed summation order): ``RelGATLoss`` -> relgat_rank_loss, ``MultiObjectiveRelLoss`` -> relgat_rank_loss +
relgat_recon_loss.  Host tensors (the CPU unit tests of the API contract) and float64 take the same formulas as
torch tensor ops.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
import torch.nn.functional as F


def _fusable(*tensors) -> bool:
    """CUDA float32 tensors go to the fused loss kernels."""
    return all(t is not None and t.is_cuda and t.dtype == torch.float32 for t in tensors)


class CosineLoss:
    @staticmethod
    def calculate(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """mean(1 - cos(pred, target)) (cosine.py:10-13); broadcasts [B, D] against [K, B, D]."""
        p = F.normalize(pred, p=2, dim=-1)
        t = F.normalize(target, p=2, dim=-1)
        return (1.0 - (p * t).sum(dim=-1)).mean()


class MSELoss:
    @staticmethod
    def calculate(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return F.mse_loss(a, b)


class RelGATLoss:
    def __init__(self, loss_type: str, self_adv_alpha: Optional[float], margin: Optional[float],
                 clamp_limit: Optional[float], run_config: Dict[str, Any]):
        self.loss_type = loss_type
        self.clamp_limit = clamp_limit
        self.margin = run_config.get("margin", margin)
        if self.margin is not None:
            self.margin = float(self.margin)
        self.self_adv_alpha = run_config.get("self_adv_alpha", self_adv_alpha)
        if self.self_adv_alpha is not None:
            self.self_adv_alpha = float(self.self_adv_alpha)
        self.use_self_adv_neg = loss_type == "self_adversarial_loss"

    def prepare_scores_and_compute_loss(self, pos_score: torch.Tensor, neg_score: torch.Tensor,
                                        sanitize: bool = False) -> torch.Tensor:
        if _fusable(pos_score, neg_score) and neg_score.dim() == 2:
            from .functional import fused_rank_loss
            return fused_rank_loss(pos_score, neg_score, "self_adversarial_loss" if self.use_self_adv_neg else "margin",
                                   self.margin, self.self_adv_alpha, sanitize)
        if sanitize:
            pos_score = torch.nan_to_num(pos_score, nan=0.0, neginf=-1e9, posinf=1e9)
            neg_score = torch.nan_to_num(neg_score, nan=0.0, neginf=-1e9, posinf=1e9)
        if self.use_self_adv_neg:
            return self._self_adversarial_loss(pos_score, neg_score)
        return self._margin_ranking_loss(pos_score, neg_score)

    def _margin_ranking_loss(self, pos_score, neg_score):
        """mean_{b,k} relu(margin + neg[b,k] - pos[b]) (relgat_loss.py:51-54)."""
        return F.relu(self.margin + neg_score - pos_score.unsqueeze(1)).mean()

    def _self_adversarial_loss(self, pos_score, neg_score):
        """-mean logsig(pos) - mean_b sum_k softmax_k(alpha*neg).detach() * logsig(-neg) (relgat_loss.py:56-71)."""
        with torch.no_grad():
            w = torch.softmax(self.self_adv_alpha * neg_score, dim=1)
        return -F.logsigmoid(pos_score).mean() - (w * F.logsigmoid(-neg_score)).sum(dim=1).mean()


class MultiObjectiveRelLoss:
    """Weighted mean of ranking + positive cosine + (1 - negative cosine) + MSE terms
    (multi_objective_loss.py:47-83); zero-weight terms leave numerator and denominator."""

    def __init__(self, *, relgat_loss: RelGATLoss, run_config: Dict[str, Any], pos_cosine_weight: float = 1.0,
                 neg_cosine_weight: float = 1.0, mse_weight: float = 0.0, relgat_weight: float = 1.0):
        self.ranking_weight = run_config.get("relgat_weight", relgat_weight)
        self.pos_cosine_weight = run_config.get("pos_cosine_weight", pos_cosine_weight)
        self.neg_cosine_weight = run_config.get("neg_cosine_weight", neg_cosine_weight)
        self.mse_weight = run_config.get("mse_weight", mse_weight)
        self.relgat_loss = relgat_loss
        self.last_recon_values = None  # (cosine_pos, cosine_neg, mse) loss values of the last fused call

    def relgat_ranking_loss(self, pos_score, neg_score):
        return self.relgat_loss.prepare_scores_and_compute_loss(pos_score=pos_score, neg_score=neg_score)

    def __call__(self, *, pos_score, neg_score, transformed_src, dst_vec, neg_dst_vec, sanitize: bool = False):
        weights = [w for w in (self.ranking_weight, self.pos_cosine_weight, self.neg_cosine_weight, self.mse_weight)
                   if w != 0.0]
        if not weights:
            raise ValueError("At least one loss weight must be non-zero.")
        self.last_recon_values = None
        if _fusable(pos_score, neg_score, transformed_src, dst_vec, neg_dst_vec) and neg_dst_vec.dim() == 3:
            from .functional import fused_recon_loss
            total, self.last_recon_values = fused_recon_loss(transformed_src, dst_vec, neg_dst_vec, self.pos_cosine_weight,
                                                             self.neg_cosine_weight, self.mse_weight)
            if self.ranking_weight != 0.0:
                total = total + self.ranking_weight * self.relgat_loss.prepare_scores_and_compute_loss(
                    pos_score=pos_score, neg_score=neg_score, sanitize=sanitize)
            return total / sum(weights)
        if sanitize:
            pos_score = torch.nan_to_num(pos_score, nan=0.0, neginf=-1e9, posinf=1e9)
            neg_score = torch.nan_to_num(neg_score, nan=0.0, neginf=-1e9, posinf=1e9)
        terms = [
            (self.ranking_weight, lambda: self.relgat_ranking_loss(pos_score, neg_score)),
            (self.pos_cosine_weight, lambda: CosineLoss.calculate(transformed_src, dst_vec)),
            (self.neg_cosine_weight, lambda: 1.0 - CosineLoss.calculate(transformed_src, neg_dst_vec)),
            (self.mse_weight, lambda: MSELoss.calculate(transformed_src, dst_vec)),
        ]
        active = [(w, fn) for w, fn in terms if w != 0.0]
        if not active:
            raise ValueError("At least one loss weight must be non-zero.")
        return torch.stack([w * fn() for w, fn in active]).sum() / sum(w for w, _ in active)


def calculate_loss(model, src_ids, rel_ids, dst_ids, pos_examples_in_batch: int, relgat_loss: RelGATLoss,
                   multi_loss: Optional[MultiObjectiveRelLoss] = None):
    """The trainer's ``_calculate_loss`` (reference trainer/relgat_projector.py:498-557 with :559-655) on the fused
    path: stack + batch-row gather -> [projection of those rows] -> score (+ transform of the positives) in one
    launch -> ranking [+ reconstruction] loss kernels.  Same return tuple:
    (pos_score [B], neg_score [B, K], loss, mse, cosine_pos, cosine_neg); the last three are 0-d tensors (the
    reference calls ``.item()`` on them for logging) or None without projection.

    The two negative layouts of the reference are reproduced as strided views of the flat score vector: without
    projection ``view(K, B).T`` (trainer:671-675), with projection ``view(B, K)`` and ``neg_dst_vec.view(B, K, D)
    .permute(1, 0, 2)`` (trainer:628-642).  Scores are sanitised like trainer:584 / 647-648."""
    b = int(pos_examples_in_batch)
    k = (int(src_ids.numel()) - b) // b if b else 0
    if not model.project_to_input_size:
        scores, _, _ = model(src_ids, rel_ids, dst_ids, transform_to_input_if_possible=False)
        pos, neg = scores[:b], scores[b:].view(k, b).t()
        loss = relgat_loss.prepare_scores_and_compute_loss(pos, neg, sanitize=True)
        return pos, neg, loss, None, None, None
    if multi_loss is None:
        raise ValueError("project_to_input_size=True needs the MultiObjectiveRelLoss (trainer:528-534)")
    scores, tr, dst_vec = model(src_ids, rel_ids, dst_ids, transform_rows=b)
    pos, neg = scores[:b], scores[b:].view(b, k)
    ndv = dst_vec[b:].view(b, k, dst_vec.size(1)).permute(1, 0, 2) if k > 0 else None
    if ndv is None:
        ndv = dst_vec.new_zeros((0, b, dst_vec.size(1)))
    loss = multi_loss(pos_score=pos, neg_score=neg, transformed_src=tr, dst_vec=dst_vec[:b], neg_dst_vec=ndv,
                      sanitize=True)
    vals = getattr(multi_loss, "last_recon_values", None)
    if vals is None:
        with torch.no_grad():
            vals = torch.stack([CosineLoss.calculate(tr, dst_vec[:b]), CosineLoss.calculate(tr, ndv),
                                MSELoss.calculate(tr, dst_vec[:b])])
    return pos, neg, loss, vals[2], vals[0], vals[1]


def fused_margin_ranking_loss(scores: torch.Tensor, num_pos: int, num_neg: int, margin: float,
                              projection_path: bool = False) -> torch.Tensor:
    """``RelGATLoss("margin")`` applied to ``split_scores(scores, ...)`` as ONE kernel on the flat CUDA
    score vector (same value; also yields d loss / d score for the scorer's backward)."""
    from .functional import fused_margin_loss
    return fused_margin_loss(scores, num_pos, num_neg, margin, projection_path)


def split_scores(scores: torch.Tensor, num_pos: int, num_neg: int, projection_path: bool = False):
    """Flat scores [B*(1+K)] (positives, then K-major negative blocks — reference
    trainer/components/relgat_batching.py:5-19) -> (pos [B], neg [B, K]).

    ``projection_path=False``: view(K, B).T (reference trainer/relgat_projector.py:657-676).
    ``projection_path=True`` : view(B, K), the pairing the reference's projection branch uses
    (trainer/relgat_projector.py:628-630)."""
    pos, flat = scores[:num_pos], scores[num_pos:]
    if projection_path:
        return pos, flat.view(num_pos, num_neg)
    return pos, flat.view(num_neg, num_pos).transpose(0, 1).contiguous()


def compute_mrr_hits(pos_score: torch.Tensor, neg_score: torch.Tensor, ks, pessimistic: bool = True):
    """MRR / Hits@k against the sampled negatives, pessimistic ties (reference core/eval.py:8-37)."""
    if pos_score.shape[0] == 0:
        return 0.0, {k: 0.0 for k in ks}
    p = torch.nan_to_num(pos_score, nan=-1e9, neginf=-1e9, posinf=1e9)
    q = torch.nan_to_num(neg_score, nan=-1e9, neginf=-1e9, posinf=1e9)
    worse = (q >= p.unsqueeze(1)) if pessimistic else (q > p.unsqueeze(1))
    ranks = 1.0 + worse.to(p.dtype).sum(dim=1)
    mrr = (1.0 / torch.clamp(ranks, min=1.0)).mean().item()
    return mrr, {k: (ranks <= float(k)).to(p.dtype).mean().item() for k in ks}
