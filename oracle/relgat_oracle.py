"""CPU oracle for the RelGAT message-passing hot path (TEST INFRASTRUCTURE ONLY).

This file restates, on the CPU, the algorithm of the reference's hot path so that
the CUDA kernels can be checked on a box where ``/root/reference`` does not exist.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package never does.

PARITY STATUS: **unpinned by the reference's own tests** — the reference ships no
tests, golden vectors or fixtures, and the segment arithmetic lives
in the un-pinned third-party ``torch_scatter`` (absent from ``/root/reference``;
restated in ``oracle/standin/torch_scatter``).  The oracle is instead pinned
against outputs of the reference's own Python files executed in the build
container: ``oracle/gen_golden.py`` writes ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks every function below against them.

Two restatements are provided:

* ``*_port`` functions — torch (fp32/fp64) code that issues the same sequence of
  tensor ops as the reference (per-head loop, materialised ``[E, F]`` gathers,
  scatter ops) and therefore reproduces its CPU results bit-for-bit and its CPU
  cost; this is what ``bench.py`` times as ``cpu_baseline`` (kind "port").
* ``*_closed`` functions — NumPy fp64 closed forms used to
  check intermediates the reference never exposes (logits, attention weights,
  analytic gradients).
"""
from __future__ import annotations

import random as _pyrandom
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.2  # reference core/model/layer.py:233
SOFTMAX_EPS = 1e-16  # reference core/model/layer.py:85 (STABLE_SOFTMAX_EPS)


# ----------------------------------------------------------------------------------
# segment primitives (the torch_scatter boundary; reference layer.py:284,290,308,316)
# ----------------------------------------------------------------------------------
def seg_sum(values: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    """``scatter_add(values, index, dim=0, dim_size=n)``: zeros + ``scatter_add_``."""
    idx = index
    while idx.dim() < values.dim():
        idx = idx.unsqueeze(-1)
    idx = idx.expand_as(values)
    out = torch.zeros((n,) + tuple(values.shape[1:]), dtype=values.dtype, device=values.device)
    return out.scatter_add_(0, idx, values)


def seg_max_const(values: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    """``scatter_max(values, index, dim=0, dim_size=n)[0]`` with empty segments = 0.

    Returned detached: the shift by the per-destination maximum has zero net
    gradient, whichever element torch_scatter routes it through.
    """
    lowest = torch.finfo(values.dtype).min
    out = torch.full((n,), lowest, dtype=values.dtype, device=values.device)
    out = out.scatter_reduce(0, index, values.detach(), reduce="amax", include_self=True)
    return torch.where(out == lowest, torch.zeros_like(out), out)


# ----------------------------------------------------------------------------------
# RelGAT layer, reference op order (reference core/model/layer.py:214-323)
# ----------------------------------------------------------------------------------
def layer_forward_port(
    x: torch.Tensor,  # [N, D_in]
    W: Sequence[torch.Tensor],  # H x [F, D_in]      (proj[h].weight, layer.py:108-110)
    A: Sequence[torch.Tensor],  # H x [R, F]         (attn_vec[h],   layer.py:113-115)
    beta: Optional[torch.Tensor],  # [R] or None     (rel_bias,      layer.py:118-121)
    edge_index: torch.Tensor,  # [2, E] int64 (row 0 = src, row 1 = dst)
    edge_type: torch.Tensor,  # [E] int64
    return_attention: bool = False,
    attn_keep: Optional[torch.Tensor] = None,  # [E, H] bool in COO edge order: keep mask of rel_attn_drop (layer.py:296-297)
    attn_p: float = 0.0,
    feat_keep: Optional[torch.Tensor] = None,  # [N, H*F] bool: keep mask of the output dropout (layer.py:321-322)
    feat_p: float = 0.0,
):
    src, dst = edge_index[0], edge_index[1]
    n = x.size(0)
    heads = len(W)
    gathered = [(x @ W[h].t())[src] for h in range(heads)]  # layer.py:220
    logits = []
    for h in range(heads):  # layer.py:226-234
        a_e = A[h][edge_type]
        z = (gathered[h] * a_e).sum(dim=-1)
        logits.append(F.leaky_relu(z, negative_slope=LEAKY_SLOPE))
    alphas = []
    for h in range(heads):  # layer.py:280-299
        m = seg_max_const(logits[h], dst, n)
        w = torch.exp(logits[h] - m[dst])
        den = seg_sum(w, dst, n).clamp_min(SOFTMAX_EPS)
        a = w / den[dst]
        if attn_keep is not None:  # F.dropout with a given mask: kept entries scaled by 1 / (1 - p)
            a = a * attn_keep[:, h].to(a.dtype) / (1.0 - attn_p)
        alphas.append(a)
    outs = [seg_sum(gathered[h] * alphas[h].unsqueeze(-1), dst, n) for h in range(heads)]  # :304-309
    if beta is not None:  # layer.py:313-318
        b = seg_sum(beta[edge_type], dst, n).unsqueeze(-1)
        outs = [o + b for o in outs]
    y = torch.cat(outs, dim=-1)  # layer.py:321 (dropout is identity at p=0 / eval)
    if feat_keep is not None:
        y = y * feat_keep.to(y.dtype) / (1.0 - feat_p)
    if return_attention:
        return y, torch.stack(logits, 1), torch.stack(alphas, 1)
    return y


def gat_stack_port(x, layers: Sequence[Dict], edge_index, edge_type, masks: Optional[Sequence[Dict]] = None):
    """``RelGATModel.single_gat_step`` without projection (reference model.py:274-287).  ``masks``: per layer the
    keyword arguments attn_keep / attn_p / feat_keep / feat_p of layer_forward_port (training-mode dropout replayed
    with given masks)."""
    for li, lp in enumerate(layers):
        x = layer_forward_port(x, lp["W"], lp["A"], lp["beta"], edge_index, edge_type, **(masks[li] if masks else {}))
        if len(layers) > 1 and li < len(layers) - 1:
            x = F.elu(x)
    return x


def projection_port(x, proj: Optional[Dict]):
    """``ProjectionHead.forward`` (reference core/model/projection.py:48-72), dropout off.

    ``proj`` = {"weights": [W0, W1, ...], "ln": [(gamma, beta), ...]} with
    ``len(ln) == len(weights) - 1``; ``None`` = no projection.
    """
    if proj is None:
        return x
    ws = proj["weights"]
    for i, w in enumerate(ws):
        x = x @ w.t()
        if i < len(ws) - 1:
            x = F.gelu(x)
            g, b = proj["ln"][i]
            x = F.layer_norm(x, (x.size(-1),), g, b)
    return x


# ----------------------------------------------------------------------------------
# scorers (reference core/scorer.py:58-94, 154-201)
# ----------------------------------------------------------------------------------
def distmult_score_port(s, rel_emb, rel_ids, t):
    return (s * rel_emb[rel_ids] * t).sum(dim=-1)  # scorer.py:80-83


def distmult_transform_port(s, rel_emb, rel_ids):
    return s * rel_emb[rel_ids]  # scorer.py:93-94


def transe_score_port(s, rel_emb, rel_ids, t, normalize: bool = True):
    r = rel_emb[rel_ids]
    if normalize:  # scorer.py:177-180 (model.py:93-95 always passes normalize=True)
        s = F.normalize(s, p=2, dim=-1)
        r = F.normalize(r, p=2, dim=-1)
        t = F.normalize(t, p=2, dim=-1)
    return -torch.norm(s + r - t, p=2, dim=-1)  # scorer.py:183-186


def transe_transform_port(s, rel_emb, rel_ids, normalize: bool = True):
    r = rel_emb[rel_ids]
    if normalize:  # scorer.py:196-200
        s = F.normalize(s, p=2, dim=-1)
        r = F.normalize(r, p=2, dim=-1)
    return s + r


def score_port(kind: str, s, rel_emb, rel_ids, t):
    if kind == "distmult":
        return distmult_score_port(s, rel_emb, rel_ids, t)
    if kind == "transe":
        return transe_score_port(s, rel_emb, rel_ids, t, True)
    raise ValueError(kind)


def transform_port(kind: str, s, rel_emb, rel_ids):
    if kind == "distmult":
        return distmult_transform_port(s, rel_emb, rel_ids)
    if kind == "transe":
        return transe_transform_port(s, rel_emb, rel_ids, True)
    raise ValueError(kind)


# ----------------------------------------------------------------------------------
# losses and metric (reference core/loss/*.py, core/eval.py, trainer :587-676)
# ----------------------------------------------------------------------------------
def split_scores_kmajor(scores: torch.Tensor, b: int, k: int):
    """No-projection path: flat negatives are K-major blocks (trainer:657-676)."""
    return scores[:b], scores[b:].view(k, b).transpose(0, 1).contiguous()


def split_scores_projection_path(scores: torch.Tensor, b: int, k: int):
    """Projection path: ``neg.view(B, K)`` on the K-major flat vector (trainer:628-630;
    SURVEY.md §B.1 — reproduces the reference's pairing, including its quirk)."""
    return scores[:b], scores[b:].view(b, k)


def margin_ranking_loss_port(pos, neg, margin: float):
    return F.relu(margin + neg - pos.unsqueeze(1).expand_as(neg)).mean()  # relgat_loss.py:51-54


def self_adversarial_loss_port(pos, neg, alpha: float):
    with torch.no_grad():  # relgat_loss.py:64-66
        w = torch.softmax(alpha * neg, dim=1)
    return -F.logsigmoid(pos).mean() - (w * F.logsigmoid(-neg)).sum(dim=1).mean()  # :68-71


def cosine_loss_port(pred, target):
    p = F.normalize(pred, p=2, dim=-1)  # cosine.py:10-13
    t = F.normalize(target, p=2, dim=-1)
    return (1.0 - (p * t).sum(dim=-1)).mean()


def multi_objective_loss_port(
    pos, neg, transformed_src, dst_vec, neg_dst_vec, *, ranking_loss,
    w_rank=1.0, w_pos=1.0, w_neg=1.0, w_mse=0.0,
):
    """``MultiObjectiveRelLoss.__call__`` (multi_objective_loss.py:47-83)."""
    parts, weights = [], []
    if w_rank != 0.0:
        weights.append(w_rank)
        parts.append(w_rank * ranking_loss(pos, neg))
    if w_pos != 0.0:
        weights.append(w_pos)
        parts.append(w_pos * cosine_loss_port(transformed_src, dst_vec))
    if w_neg != 0.0:
        weights.append(w_neg)
        parts.append(w_neg * (1.0 - cosine_loss_port(transformed_src, neg_dst_vec)))
    if w_mse != 0.0:
        weights.append(w_mse)
        parts.append(w_mse * F.mse_loss(transformed_src, dst_vec))
    if not parts:
        raise ValueError("At least one loss weight must be non-zero.")
    return torch.stack(parts).sum() / sum(weights)


def mrr_hits_port(pos, neg, ks: Sequence[int]):
    """``RelgatEval.compute_mrr_hits`` pessimistic ties (core/eval.py:8-37)."""
    if pos.shape[0] == 0:
        return 0.0, {k: 0.0 for k in ks}
    p = torch.nan_to_num(pos, nan=-1e9, neginf=-1e9, posinf=1e9)
    q = torch.nan_to_num(neg, nan=-1e9, neginf=-1e9, posinf=1e9)
    ranks = 1.0 + (q >= p.unsqueeze(1)).to(p.dtype).sum(dim=1)
    mrr = (1.0 / torch.clamp(ranks, min=1.0)).mean().item()
    return mrr, {k: (ranks <= float(k)).to(p.dtype).mean().item() for k in ks}


# ----------------------------------------------------------------------------------
# full training step, reference op order (model.py:99-142 + trainer:498-557)
# ----------------------------------------------------------------------------------
def train_step_port(
    x0, layers, rel_emb, edge_index, edge_type, src_ids, rel_ids, dst_ids, *,
    scorer: str = "distmult", b: int, k: int, margin: float = 1.0,
    loss_type: str = "margin", self_adv_alpha: float = 1.0,
    proj: Optional[Dict] = None, weights=(1.0, 1.0, 1.0, 0.0),
):
    """Loss of one step; call ``.backward()`` on the result for the gradients."""
    x = gat_stack_port(x0, layers, edge_index, edge_type)

    def rank_loss(p, n):
        if loss_type == "margin":
            return margin_ranking_loss_port(p, n, margin)
        return self_adversarial_loss_port(p, n, self_adv_alpha)

    if proj is None:  # trainer:510-521
        scores = score_port(scorer, x[src_ids], rel_emb, rel_ids, x[dst_ids])
        pos, neg = split_scores_kmajor(scores, b, k)
        return rank_loss(pos, neg), pos, neg
    x = projection_port(x, proj)  # model.py:289-290, trainer:608
    ps, pd = x[src_ids[:b]], x[dst_ids[:b]]
    pos = score_port(scorer, ps, rel_emb, rel_ids[:b], pd)
    tr = transform_port(scorer, ps, rel_emb, rel_ids[:b])
    nd = x[dst_ids[b:]]
    neg = score_port(scorer, x[src_ids[b:]], rel_emb, rel_ids[b:], nd).view(b, k)
    ndv = nd.view(b, k, tr.shape[1]).permute(1, 0, 2).contiguous()  # trainer:634-642
    pos = torch.nan_to_num(pos, nan=0.0, neginf=-1e9, posinf=1e9)
    neg = torch.nan_to_num(neg, nan=0.0, neginf=-1e9, posinf=1e9)
    loss = multi_objective_loss_port(
        pos, neg, tr, pd, ndv, ranking_loss=rank_loss,
        w_rank=weights[0], w_pos=weights[1], w_neg=weights[2], w_mse=weights[3],
    )
    return loss, pos, neg


# ----------------------------------------------------------------------------------
# integer structures: CSR / CSC / by-relation / owner buckets
# The reference is COO-only (dataset/relgat_dataset.py:123-137); "bit-exact CSR" is
# defined as the stable by-destination ordering of that COO.
# ----------------------------------------------------------------------------------
def csr_by_key_np(key: np.ndarray, n_keys: int):
    """Stable counting order of ``key``: returns (ptr[n_keys+1] int32, perm[E] int32)."""
    key = np.asarray(key, dtype=np.int64)
    perm = np.argsort(key, kind="stable").astype(np.int32)
    counts = np.bincount(key, minlength=n_keys).astype(np.int64)
    ptr = np.zeros(n_keys + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    return ptr.astype(np.int32), perm


def graph_index_np(src: np.ndarray, dst: np.ndarray, rel: np.ndarray, n: int, r: int) -> Dict[str, np.ndarray]:
    """All integer structures the kernels consume, from the reference's COO order.

    csr_*  : edges in stable by-destination order (slot s <-> original edge csr_perm[s])
    csc_*  : stable by-source order; csc_slot[t] = CSR slot of the t-th by-source edge
    rel_*  : stable by-relation order of CSR slots
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    rel = np.asarray(rel, dtype=np.int64)
    rowptr, perm = csr_by_key_np(dst, n)
    col_src = src[perm].astype(np.int32)
    col_rel = rel[perm].astype(np.int32)
    col_dst = dst[perm].astype(np.int32)
    colptr, cperm = csr_by_key_np(col_src, n)  # order CSR slots by source, stable
    relptr, rperm = csr_by_key_np(col_rel, r)
    return {
        "rowptr": rowptr, "csr_perm": perm, "csr_src": col_src, "csr_rel": col_rel, "csr_dst": col_dst,
        "colptr": colptr, "csc_slot": cperm, "csc_dst": col_dst[cperm], "csc_rel": col_rel[cperm],
        "relptr": relptr, "rel_slot": rperm,
    }


def partition_bounds_np(dst: np.ndarray, n: int, world: int, balance: str = "nodes") -> np.ndarray:
    """Contiguous destination ranges [b_g, b_{g+1}) per rank."""
    if balance == "nodes":
        return np.array([(n * g) // world for g in range(world + 1)], dtype=np.int64)
    counts = np.bincount(np.asarray(dst, dtype=np.int64), minlength=n)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = cum[-1]
    b = [0]
    for g in range(1, world):
        b.append(int(np.searchsorted(cum, (total * g + world - 1) // world, side="left")))
    b.append(n)
    return np.maximum.accumulate(np.array(b, dtype=np.int64))


def bucket_edges_np(src, dst, rel, bounds: np.ndarray):
    """Stable bucketing of COO edges by owner(dst).  Returns per-rank (src, dst, rel, eid)."""
    dst = np.asarray(dst, dtype=np.int64)
    owner = np.searchsorted(bounds, dst, side="right") - 1
    out = []
    for g in range(len(bounds) - 1):
        eid = np.nonzero(owner == g)[0]
        out.append((np.asarray(src)[eid], dst[eid], np.asarray(rel)[eid], eid.astype(np.int64)))
    return out


# ----------------------------------------------------------------------------------
# closed forms in fp64
# ----------------------------------------------------------------------------------
def layer_forward_closed(P: np.ndarray, A: np.ndarray, beta: np.ndarray, g: Dict[str, np.ndarray]):
    """P [N,H,F], A [H,R,F], beta [R]; returns out [N,H,F], z [E,H], alpha [E,H], bias [N]
    with edges in CSR slot order."""
    P = np.asarray(P, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    n, h, f = P.shape
    rowptr, cs, cr = g["rowptr"], g["csr_src"], g["csr_rel"]
    e = len(cs)
    z = np.einsum("ehf,hef->eh", P[cs], A[:, cr, :]) if e else np.zeros((0, h))
    eps = np.where(z > 0, z, LEAKY_SLOPE * z)
    alpha = np.zeros_like(eps)
    out = np.zeros((n, h, f))
    bias = np.zeros(n)
    for j in range(n):
        lo, hi = int(rowptr[j]), int(rowptr[j + 1])
        if hi == lo:
            continue
        w = np.exp(eps[lo:hi] - eps[lo:hi].max(axis=0, keepdims=True))
        a = w / np.maximum(w.sum(axis=0, keepdims=True), SOFTMAX_EPS)
        alpha[lo:hi] = a
        bias[j] = beta[cr[lo:hi]].sum()
        out[j] = np.einsum("eh,ehf->hf", a, P[cs[lo:hi]]) + bias[j]
    return out, z, alpha, bias


def layer_backward_closed(G, P, A, g, z, alpha):
    """Returns dP [N,H,F], dA [H,R,F], dbeta [R], dz [E,H] (CSR slot order)."""
    G = np.asarray(G, dtype=np.float64)
    P = np.asarray(P, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    cs, cr, cd = g["csr_src"], g["csr_rel"], g["csr_dst"]
    n, h, f = P.shape
    r = A.shape[1]
    dalpha = np.einsum("ehf,ehf->eh", G[cd], P[cs])
    t = np.zeros((n, h))
    np.add.at(t, cd, alpha * dalpha)
    deps = alpha * (dalpha - t[cd])
    dz = deps * np.where(z > 0, 1.0, LEAKY_SLOPE)
    dP = np.zeros_like(P)
    np.add.at(dP, cs, alpha[:, :, None] * G[cd] + dz[:, :, None] * np.transpose(A[:, cr, :], (1, 0, 2)))
    dA = np.zeros_like(A)
    contrib = dz[:, :, None] * P[cs]  # [E,H,F]
    for hh in range(h):
        np.add.at(dA[hh], cr, contrib[:, hh, :])
    dbeta = np.zeros(r)
    np.add.at(dbeta, cr, G[cd].sum(axis=(1, 2)))
    return dP, dA, dbeta, dz


# ----------------------------------------------------------------------------------
# seeded negative sampling (reference dataset/edge.py:71-115) — consumes the global
# CPython ``random`` stream exactly like EdgeDataset.__getitem__ does.
# ----------------------------------------------------------------------------------
def sample_batch_port(edges: Sequence[Tuple[int, int, int]], idxs: Sequence[int], n_nodes: int, k: int,
                      rng=_pyrandom):
    """Returns flat (src, rel, dst) int64 arrays of length B*(1+K): positives first, then
    K-major negative blocks (trainer/components/relgat_batching.py:5-19)."""
    ids = list(range(n_nodes))  # dataset/relgat_dataset.py:97 (all_node_ids = range(N))
    b = len(idxs)
    src = np.empty(b * (1 + k), dtype=np.int64)
    rel = np.empty_like(src)
    dst = np.empty_like(src)
    for i, ei in enumerate(idxs):
        s, d, r = edges[ei]
        src[i], rel[i], dst[i] = s, r, d
        for kk in range(k):
            c = rng.choice(ids)
            while c == d:
                c = rng.choice(ids)
            o = b + kk * b + i
            src[o], rel[o], dst[o] = s, r, c
    return src, rel, dst
