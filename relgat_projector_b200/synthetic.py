"""Seeded synthetic knowledge graphs of the shapes named in BASELINE.json.

Small graphs are produced as the reference's three input objects (``node2emb``, ``rel2idx``,
``edge_index_raw``) so they can go through the reference's own dataset pipeline; large ones
are generated directly as tensors (the Python-object form of 50 M triples does not fit a test).
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

CONFIGS = {
    # name: N, T (triplets), R, D_in, L, H, F, scorer, B, K, projection
    "c1": dict(N=10_000, T=50_000, R=50, D_in=1024, L=2, H=4, F=200, scorer="distmult", B=256, K=4, proj=False),
    "c2": dict(N=300_000, T=1_500_000, R=50, D_in=1024, L=2, H=4, F=200, scorer="distmult", B=1024, K=4, proj=False),
    "c3": dict(N=300_000, T=1_500_000, R=50, D_in=1024, L=3, H=8, F=200, scorer="transe", B=4096, K=4, proj=True),
    "c4": dict(N=5_000_000, T=50_000_000, R=200, D_in=768, L=2, H=4, F=200, scorer="distmult", B=1024, K=4, proj=False),
    "tiny": dict(N=2_000, T=12_000, R=12, D_in=64, L=2, H=4, F=40, scorer="distmult", B=64, K=4, proj=False),
}


def reference_inputs(n: int, t: int, r: int, d_in: int, seed: int = 42, skew: float = 0.0):
    """(node2emb, rel2idx, edge_index_raw) exactly as the reference handler would load them
    (handlers/models/relgat.py:12-41).  ``skew`` > 0 draws destinations and relations from a
    Zipf-like law to exercise long segments."""
    rng = np.random.default_rng(seed)
    node2emb = {i: rng.standard_normal(d_in).astype(np.float32) for i in range(n)}
    rel2idx = {f"rel_{i}": i for i in range(r)}
    src, dst, rel = _triples(rng, n, t, r, skew)
    raw = [(int(a), int(b), f"rel_{int(c)}") for a, b, c in zip(src, dst, rel)]
    return node2emb, rel2idx, raw


def _triples(rng, n, t, r, skew):
    src = rng.integers(0, n, size=t)
    if skew > 0:
        pn = 1.0 / np.arange(1, n + 1) ** skew
        pr = 1.0 / np.arange(1, r + 1) ** skew
        dst = rng.choice(n, size=t, p=pn / pn.sum())
        rel = rng.choice(r, size=t, p=pr / pr.sum())
    else:
        dst = rng.integers(0, n, size=t)
        rel = rng.integers(0, r, size=t)
    clash = src == dst
    while clash.any():  # resample self-pairs
        src[clash] = rng.integers(0, n, size=int(clash.sum()))
        clash = src == dst
    return src.astype(np.int64), dst.astype(np.int64), rel.astype(np.int64)


@dataclass
class TensorKG:
    node_emb: torch.Tensor       # [N, D_in] fp32
    edge_index: torch.Tensor     # [2, E] int64 (train split = message-passing graph)
    edge_type: torch.Tensor      # [E] int64
    train_triples: torch.Tensor  # [E, 3] int64 (src, dst, rel) — same edges, batch source
    eval_triples: torch.Tensor   # [T-E, 3]
    num_rel: int


def tensor_kg(n: int, t: int, r: int, d_in: int, seed: int = 42, train_ratio: float = 0.9,
              device: str = "cpu", skew: float = 0.0, emb_on_device: bool = True,
              locality: float = 0.0, blocks: int = 1) -> TensorKG:
    """Large-graph variant: same distributions, shuffled with a seeded permutation and split
    int(train_ratio*T) / rest like reference dataset/relgat_dataset.py:70-88.
    ``locality`` > 0 (multi-GPU supplementary runs only): that fraction of the triples gets its head
    redrawn from the tail's block of ``blocks`` equal node ranges — a graph whose partition cuts few edges,
    as a partitioner achieves on real KGs; 0 = uniformly random heads, the worst case for a partition."""
    rng = np.random.default_rng(seed)
    src, dst, rel = _triples(rng, n, t, r, skew)
    if locality > 0 and blocks > 1:
        size = -(-n // blocks)
        pick = rng.random(t) < locality
        lo = (dst // size) * size
        hi = np.minimum(lo + size, n)
        near = lo + (rng.random(t) * (hi - lo)).astype(np.int64)
        near = np.where(near == dst, np.where(near + 1 < hi, near + 1, lo), near)  # no self-pairs
        src = np.where(pick, near, src)
    perm = rng.permutation(t)
    src, dst, rel = src[perm], dst[perm], rel[perm]
    n_train = int(train_ratio * t)
    trip = torch.from_numpy(np.stack([src, dst, rel], 1))
    dev = torch.device(device)
    if emb_on_device and dev.type == "cuda":
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        emb = torch.randn((n, d_in), generator=g, device=dev, dtype=torch.float32)
    else:
        emb = torch.from_numpy(rng.standard_normal((n, d_in), dtype=np.float32)).to(dev)
    train = trip[:n_train].to(dev)
    return TensorKG(node_emb=emb, edge_index=train[:, :2].t().contiguous(), edge_type=train[:, 2].contiguous(),
                    train_triples=train, eval_triples=trip[n_train:].to(dev), num_rel=r)


def sample_batch(train_triples: torch.Tensor, n_nodes: int, b: int, k: int, generator: torch.Generator):
    """Vectorised corrupted-tail batch in the reference's flat layout (positives, then K-major
    negative blocks; trainer/components/relgat_batching.py:5-19).  Used by the benchmark; the
    bit-exact CPython-stream sampler is ``batching.ReferenceStreamSampler``."""
    dev = train_triples.device
    idx = torch.randint(0, train_triples.size(0), (b,), generator=generator, device=generator.device).to(dev)
    pos = train_triples[idx]
    s, d, r = pos[:, 0], pos[:, 1], pos[:, 2]
    neg = torch.randint(0, n_nodes - 1, (k, b), generator=generator, device=generator.device).to(dev)
    neg = neg + (neg >= d.unsqueeze(0)).to(neg.dtype)  # uniform over nodes != dst, no rejection loop
    src_ids = s.repeat(k + 1)
    rel_ids = r.repeat(k + 1)
    dst_ids = torch.cat([d, neg.reshape(-1)])
    return src_ids, rel_ids, dst_ids
