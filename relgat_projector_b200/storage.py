"""On-disk node table for the RelGAT inputs.

The reference loads node embeddings as a pickled ``dict[int, vector]`` and rebuilds the ``[N, D_in]`` matrix with a
per-node ``torch.as_tensor`` + ``torch.stack`` on every start (reference handlers/models/relgat.py:12-20,
dataset/relgat_dataset.py:61-68).  Here the same matrix — rows in ascending node-id order, exactly the reference's
``sorted(node2emb.keys())`` — is written once as two ``.npy`` files and memory-mapped afterwards: start-up cost is
independent of N, and the host copy that feeds the GPU is the page cache itself."""
from __future__ import annotations

import pickle
from typing import Dict, Mapping, Tuple

import numpy as np
import torch


def write_node_table(node2emb: Mapping[int, object], prefix: str) -> Tuple[int, int]:
    """Writes ``<prefix>.ids.npy`` (int64 [N], ascending) and ``<prefix>.emb.npy`` (float32 [N, D]); returns (N, D).
    Rows are streamed into a memory-mapped file, so the dict's vectors are never stacked in memory."""
    ids = np.array(sorted(int(k) for k in node2emb.keys()), dtype=np.int64)
    if ids.size == 0:
        raise ValueError("node2emb is empty")
    keys = {int(k): k for k in node2emb.keys()}
    first = np.asarray(torch.as_tensor(node2emb[keys[int(ids[0])]]).to(torch.float32).numpy())
    if first.ndim != 1:
        raise ValueError("node vectors must be one-dimensional")
    out = np.lib.format.open_memmap(prefix + ".emb.npy", mode="w+", dtype=np.float32, shape=(ids.size, first.size))
    for row, nid in enumerate(ids):
        v = torch.as_tensor(node2emb[keys[int(nid)]]).to(torch.float32).numpy()
        if v.shape != first.shape:
            raise ValueError(f"node {int(nid)}: vector of shape {v.shape}, expected {first.shape}")
        out[row] = v
    out.flush()
    del out
    np.save(prefix + ".ids.npy", ids)
    return int(ids.size), int(first.size)


def load_node_table(prefix: str, mmap: bool = True) -> Tuple[np.ndarray, torch.Tensor]:
    """(ids int64 [N], node_emb float32 [N, D]).  With ``mmap`` the tensor is a copy-on-write view of the file:
    nothing is read until rows are touched (e.g. by ``.to(device)``), and the file is never modified."""
    ids = np.load(prefix + ".ids.npy")
    emb = np.load(prefix + ".emb.npy", mmap_mode="c" if mmap else None)
    if emb.ndim != 2 or emb.shape[0] != ids.shape[0] or emb.dtype != np.float32:
        raise ValueError("node table files do not match (rows / dtype)")
    return ids, torch.from_numpy(emb)


def id_to_row(ids: np.ndarray) -> Dict[int, int]:
    """The reference's ``id2idx`` (dataset/relgat_dataset.py:63)."""
    return {int(nid): row for row, nid in enumerate(ids)}


def convert_pickled_nodes(pickle_path: str, prefix: str) -> Tuple[int, int]:
    """One-off conversion of the reference's pickle-of-dict embedding file."""
    with open(pickle_path, "rb") as f:
        node2emb = pickle.load(f)
    return write_node_table(node2emb, prefix)
