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


# ---------------------------------------------------------------------------------------------
# checkpoints without the frozen node table
# ---------------------------------------------------------------------------------------------
FROZEN_KEY = "node_emb_fixed"  # reference core/model/model.py:32: a BUFFER, so torch.save(state_dict()) stores all N rows


def _table_fingerprint(t: torch.Tensor) -> Dict[str, object]:
    """Shape, dtype and a cheap content checksum (sum of the table and of every 1009th row) of the frozen embeddings."""
    x = t.detach()
    probe = x[:: max(1, 1009)].double()
    return {"shape": list(x.shape), "dtype": str(x.dtype), "sum": float(x.double().sum()), "probe_sum": float(probe.sum())}


def save_trainable_state(model: torch.nn.Module, path: str) -> int:
    """``torch.save`` of the model's state dict WITHOUT the frozen ``node_emb_fixed`` buffer (the reference writes the
    whole [N, D_in] input matrix into every checkpoint: handlers/storage.py:45-56 — 1.2 GB per checkpoint at config 2
    for 4.6 M trainable parameters).  A fingerprint of the table is stored instead and verified on load.  Returns the
    number of bytes written."""
    import os
    sd = model.state_dict()
    frozen = sd.pop(FROZEN_KEY, None)
    payload = {"state_dict": sd, "frozen": None if frozen is None else _table_fingerprint(frozen), "format": 1}
    torch.save(payload, path)
    return os.path.getsize(path)


def load_trainable_state(model: torch.nn.Module, path: str, map_location=None, check_table: bool = True) -> None:
    """Loads a ``save_trainable_state`` file into a model that was constructed with the same frozen node table (e.g.
    from ``load_node_table``).  Every key but the frozen buffer must match (strict); the table's fingerprint is compared
    unless ``check_table`` is off.  A full reference-style state dict (with the buffer) is accepted as well."""
    payload = torch.load(path, map_location=map_location)
    if isinstance(payload, dict) and "state_dict" in payload and "format" in payload:
        sd, fp = payload["state_dict"], payload.get("frozen")
    else:  # a plain state dict as the reference saves it
        sd, fp = dict(payload), None
        sd.pop(FROZEN_KEY, None)
    own = model.state_dict()
    missing = [k for k in own if k != FROZEN_KEY and k not in sd]
    unexpected = [k for k in sd if k not in own]
    if missing or unexpected:
        raise KeyError(f"checkpoint does not match the model: missing {missing}, unexpected {unexpected}")
    if check_table and fp is not None and FROZEN_KEY in own:
        have = _table_fingerprint(own[FROZEN_KEY])
        if have["shape"] != fp["shape"] or abs(have["sum"] - fp["sum"]) > 1e-6 * max(1.0, abs(fp["sum"])) \
                or abs(have["probe_sum"] - fp["probe_sum"]) > 1e-6 * max(1.0, abs(fp["probe_sum"])):
            raise ValueError("the model's frozen node table is not the one this checkpoint was trained on")
    model.load_state_dict(sd, strict=False)
