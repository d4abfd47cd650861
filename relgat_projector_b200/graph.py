"""Device-resident graph index for the RelGAT kernels.

The reference hands every layer call the COO message-passing graph ``edge_index [2, E]`` /
``edge_type [E]`` as int64 tensors (reference dataset/relgat_dataset.py:123-137,
core/model/layer.py:131-136).  The kernels consume the stable by-destination (CSR), by-source
(CSC) and by-relation orderings of that COO, built once per graph by
``relgat_graph_index_build`` and cached per (edge_index, edge_type) pair.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _lib

REL_CHUNK = 256  # edges per by-relation chunk of the dA / dbeta pass
STREAM_CHUNK_EDGES = 64  # target edges per warp-chunk of the forward / by-source streaming kernels
STREAM_CHUNK_NODES = 64  # hard cap on segments per chunk (the kernels hold the pointer window in 3 registers)


def stream_chunks(ptr: torch.Tensor, chunk_edges: int = STREAM_CHUNK_EDGES,
                  chunk_nodes: int = STREAM_CHUNK_NODES) -> torch.Tensor:
    """Cut a CSR pointer array [n+1] into chunks at segment boundaries: a new chunk starts at
    segment j when its first edge crosses a multiple of ``chunk_edges`` or j is a multiple of
    ``chunk_nodes``.  Returns chunk_node [n_chunks+1] (int32).  Deterministic, one-off per graph."""
    n = ptr.numel() - 1
    if n <= 0:
        return torch.zeros(1, dtype=torch.int32, device=ptr.device)
    start = ptr[:-1].to(torch.int64)
    idx = torch.arange(n, device=ptr.device, dtype=torch.int64)
    key = (start // chunk_edges) * (n // chunk_nodes + 2) + idx // chunk_nodes
    new = torch.ones(n, dtype=torch.bool, device=ptr.device)
    new[1:] = key[1:] != key[:-1]
    heads = torch.nonzero(new).flatten()
    return torch.cat([heads, torch.tensor([n], device=ptr.device)]).to(torch.int32)


class GraphIndex:
    """All integer structures of one message-passing graph, as int32 CUDA tensors."""

    def __init__(self, edge_index: torch.Tensor, edge_type: torch.Tensor, num_nodes: int, num_rel: int,
                 validate: bool = True, num_src_nodes: Optional[int] = None):
        """``num_nodes`` = destination rows (segments); ``num_src_nodes`` = rows of the feature matrix
        the sources index (defaults to ``num_nodes``; larger on a destination-range partition)."""
        _lib.require_cuda(edge_index, edge_type)
        if edge_index.dim() != 2 or edge_index.size(0) != 2:
            raise ValueError("edge_index must have shape [2, E]")
        if edge_type.dim() != 1 or edge_type.numel() != edge_index.size(1):
            raise ValueError("edge_type must have shape [E]")
        if edge_index.dtype != torch.int64 or edge_type.dtype != torch.int64:
            raise TypeError("edge_index / edge_type must be int64 (as produced by the reference dataset)")
        dev = edge_index.device
        E = int(edge_index.size(1))
        N, R = int(num_nodes), int(num_rel)
        NS = int(num_src_nodes) if num_src_nodes is not None else N
        if validate and E > 0:
            lo = int(min(edge_index.min().item(), edge_type.min().item()))
            if (lo < 0 or int(edge_index[0].max().item()) >= NS or int(edge_index[1].max().item()) >= N
                    or int(edge_type.max().item()) >= R):
                raise IndexError("edge_index / edge_type out of range for num_nodes / num_rel")
        self.N, self.N_src, self.E, self.R, self.device = N, NS, E, R, dev
        src = edge_index[0].contiguous()
        dst = edge_index[1].contiguous()
        rel = edge_type.contiguous()
        i32 = dict(dtype=torch.int32, device=dev)
        self.rowptr = torch.empty(N + 1, **i32)
        self.colptr = torch.empty(NS + 1, **i32)
        self.relptr = torch.empty(R + 1, **i32)
        names = ["csr_perm", "csr_src", "csr_rel", "csr_dst", "csc_slot", "csc_dst", "csc_rel", "rel_slot"]
        for n in names:
            setattr(self, n, torch.empty(max(E, 1), **i32)[:E])
        lib = _lib.load()
        ws_bytes = int(lib.relgat_graph_index_workspace_bytes(E))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.relgat_graph_index_build(
                _lib.ptr(src), _lib.ptr(dst), _lib.ptr(rel), E, N, NS, R,
                _lib.ptr(self.rowptr), _lib.ptr(self.csr_perm), _lib.ptr(self.csr_src), _lib.ptr(self.csr_rel),
                _lib.ptr(self.csr_dst), _lib.ptr(self.colptr), _lib.ptr(self.csc_slot), _lib.ptr(self.csc_dst),
                _lib.ptr(self.csc_rel), _lib.ptr(self.relptr), _lib.ptr(self.rel_slot),
                _lib.ptr(ws), ws_bytes, stream)
        _lib.check(rc, "relgat_graph_index_build")
        self._build_rel_chunks()
        self.fwd_chunk_node = stream_chunks(self.rowptr)
        self.src_chunk_node = stream_chunks(self.colptr)
        self.n_fwd_chunks = int(self.fwd_chunk_node.numel()) - 1
        self.n_src_chunks = int(self.src_chunk_node.numel()) - 1
        self.max_in_degree = int((self.rowptr[1:] - self.rowptr[:-1]).max().item()) if N > 0 else 0
        self.max_out_degree = int((self.colptr[1:] - self.colptr[:-1]).max().item()) if NS > 0 else 0
        del ws

    def _build_rel_chunks(self) -> None:
        """Fixed-size chunks of the by-relation order; a chunk never spans two relations."""
        relptr = self.relptr.cpu().tolist()  # R+1 ints, one-off
        lo, hi, cptr = [], [], [0]
        for r in range(self.R):
            a, b = relptr[r], relptr[r + 1]
            for s in range(a, b, REL_CHUNK):
                lo.append(s)
                hi.append(min(b, s + REL_CHUNK))
            cptr.append(len(lo))
        i32 = dict(dtype=torch.int32, device=self.device)
        self.n_chunks = len(lo)
        self.chunk_lo = torch.tensor(lo if lo else [0], **i32)[: self.n_chunks]
        self.chunk_hi = torch.tensor(hi if hi else [0], **i32)[: self.n_chunks]
        self.rel_chunk_ptr = torch.tensor(cptr, **i32)

    def as_numpy(self) -> Dict[str, "object"]:
        keys = ["rowptr", "csr_perm", "csr_src", "csr_rel", "csr_dst", "colptr", "csc_slot", "csc_dst",
                "csc_rel", "relptr", "rel_slot"]
        return {k: getattr(self, k).cpu().numpy() for k in keys}


_CACHE: Dict[Tuple, GraphIndex] = {}
_CACHE_MAX = 8


def get_graph_index(edge_index: torch.Tensor, edge_type: torch.Tensor, num_nodes: int, num_rel: int) -> GraphIndex:
    """Cached GraphIndex for the tensors the reference passes to every ``RelGATLayer.forward``.

    Keyed on storage identity + version counters, so an in-place edit of the COO tensors
    rebuilds the index.
    """
    key = (edge_index.data_ptr(), edge_type.data_ptr(), edge_index._version, edge_type._version,
           tuple(edge_index.shape), int(num_nodes), int(num_rel), str(edge_index.device))
    g = _CACHE.get(key)
    if g is None:
        if len(_CACHE) >= _CACHE_MAX:
            _CACHE.pop(next(iter(_CACHE)))
        g = GraphIndex(edge_index, edge_type, num_nodes, num_rel)
        g._keepalive = (edge_index, edge_type)  # pins the storages so the pointer key cannot be recycled
        _CACHE[key] = g
    return g


def clear_cache() -> None:
    _CACHE.clear()
