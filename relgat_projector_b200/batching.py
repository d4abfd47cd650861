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


class NativeStreamSampler(ReferenceStreamSampler):
    """Same stream, same outputs as ``ReferenceStreamSampler``, but the B*K ``random.choice`` draws run in
    C (``relgat_host_sample_batch``): CPython's MT19937 state is handed over with ``random.getstate()``
    and handed back with ``random.setstate()``, so any other user of ``random`` is unaffected."""

    def __init__(self, edges, num_nodes: int, num_neg: int, batch_size: int, shuffle: bool = True):
        super().__init__(edges, num_nodes, num_neg, batch_size, shuffle)
        self._edges_np = np.ascontiguousarray(np.asarray(edges, dtype=np.int64).reshape(-1, 3))

    def build(self, idxs):
        from . import _lib
        lib = _lib.load()
        version, internal, gauss = random.getstate()
        state = np.array(internal, dtype=np.uint32)
        idx = np.ascontiguousarray(np.asarray(idxs, dtype=np.int64))
        b, k = int(idx.size), self.num_neg
        out = np.empty((3, b * (1 + k)), dtype=np.int64)
        rc = lib.relgat_host_sample_batch(
            state.ctypes.data, self._edges_np.ctypes.data, self._edges_np.shape[0], idx.ctypes.data, b, k,
            self.num_nodes, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data)
        _lib.check(rc, "relgat_host_sample_batch")
        random.setstate((version, tuple(int(v) for v in state), gauss))
        return torch.from_numpy(out[0]), torch.from_numpy(out[1]), torch.from_numpy(out[2])


def native_shuffle_and_split(edge_index_raw, train_ratio: float):
    """``shuffle_and_split`` with the permutation drawn in C from CPython's own generator state."""
    from . import _lib
    lib = _lib.load()
    version, internal, gauss = random.getstate()
    state = np.array(internal, dtype=np.uint32)
    perm = np.arange(len(edge_index_raw), dtype=np.int64)
    _lib.check(lib.relgat_host_shuffle(state.ctypes.data, perm.ctypes.data, perm.size), "relgat_host_shuffle")
    random.setstate((version, tuple(int(v) for v in state), gauss))
    shuffled = [edge_index_raw[i] for i in perm]
    edge_index_raw[:] = shuffled  # in place, like random.shuffle
    n_train = int(train_ratio * len(edge_index_raw))
    return edge_index_raw[:n_train], edge_index_raw[n_train:]
