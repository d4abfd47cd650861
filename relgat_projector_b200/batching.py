"""Host-side batch construction that reproduces the reference's RNG stream bit for bit
(reference dataset/edge.py:71-115, dataset/relgat_dataset.py:70-121,
trainer/components/relgat_batching.py:5-19).

The reference draws negatives with CPython's ``random.choice`` (one draw per negative, redraw
while it equals the true tail) inside ``EdgeDataset.__getitem__`` and orders batches with torch's
``DataLoader(shuffle=True)``.  ``ReferenceStreamSampler`` consumes both streams in the same order
but builds the three flat int64 id vectors directly instead of 3*(1+K) one-element tensors per
positive.
"""
from __future__ import annotations

import random
from typing import Iterator, List, Sequence, Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader


def shuffle_and_split(edge_index_raw: List[Tuple[int, int, object]], train_ratio: float):
    """In-place ``random.shuffle`` then int(train_ratio * T) split (relgat_dataset.py:70-88)."""
    random.shuffle(edge_index_raw)
    n_train = int(train_ratio * len(edge_index_raw))
    return edge_index_raw[:n_train], edge_index_raw[n_train:]


class ReferenceStreamSampler:
    def __init__(self, edges: Sequence[Tuple[int, int, int]], num_nodes: int, num_neg: int, batch_size: int,
                 shuffle: bool = True):
        self.edges = edges
        self.num_nodes = int(num_nodes)
        self.num_neg = int(num_neg)
        self.node_ids = list(range(self.num_nodes))  # all_node_ids = range(N) (relgat_dataset.py:97)
        # torch's own DataLoader orders the indices, so its generator is consumed identically
        self._loader = DataLoader(range(len(edges)), batch_size=batch_size, shuffle=shuffle, num_workers=0,
                                  collate_fn=list)

    def __len__(self) -> int:
        return len(self._loader)

    def build(self, idxs: Sequence[int]):
        b, k = len(idxs), self.num_neg
        src = np.empty(b * (1 + k), dtype=np.int64)
        rel = np.empty_like(src)
        dst = np.empty_like(src)
        choice = random.choice
        ids = self.node_ids
        for i, ei in enumerate(idxs):
            s, d, r = self.edges[ei]
            src[i], rel[i], dst[i] = s, r, d
            for kk in range(k):
                c = choice(ids)
                while c == d:
                    c = choice(ids)
                o = b + kk * b + i  # K-major negative blocks
                src[o], rel[o], dst[o] = s, r, c
        return torch.from_numpy(src), torch.from_numpy(rel), torch.from_numpy(dst)

    def __iter__(self) -> Iterator:
        for idxs in self._loader:
            yield self.build(idxs)
