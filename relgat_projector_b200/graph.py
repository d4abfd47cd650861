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
import os as _os

# target edges per warp-chunk of the forward / by-source streaming kernels (env override: experiments)
STREAM_CHUNK_EDGES = int(_os.environ.get("RELGAT_CHUNK_EDGES", "32"))
STREAM_CHUNK_NODES = 64  # hard cap on segments per chunk (the kernels hold the pointer window in 3 registers)


LONG_SEGMENT = 512   # segments with more edges are split across warps (heavy-tailed graphs)
PART_EDGES = 256     # edges per part of a split segment


class StreamChunks:
    """Work decomposition of a CSR pointer array for the streaming edge kernels.

    ``chunks`` int32 [n_chunks, 4] = (first segment, number of segments (1..64), part slot or -1, 0):
    ordinary chunks cover whole segments, ~``chunk_edges`` edges; a segment longer than
    ``long_segment`` is isolated and cut into parts of ``part_edges`` edges, one chunk per part,
    whose partial results a merge kernel combines in part order (deterministic).
    ``parts`` int32 [n_parts, 2] = (first edge, end edge); ``long_node`` [n_long] and
    ``long_part_ptr`` [n_long+1] list the split segments and their part ranges.
    """

    def __init__(self, ptr: torch.Tensor, chunk_edges: int = STREAM_CHUNK_EDGES,
                 chunk_nodes: int = STREAM_CHUNK_NODES, long_segment: int = LONG_SEGMENT,
                 part_edges: int = PART_EDGES):
        dev = ptr.device
        n = ptr.numel() - 1
        i32 = dict(dtype=torch.int32, device=dev)
        if n <= 0:
            self.chunks = torch.zeros((0, 4), **i32)
            self.parts = torch.zeros((0, 2), **i32)
            self.long_node = torch.zeros((0,), **i32)
            self.long_part_ptr = torch.zeros((1,), **i32)
            self.n_chunks = self.n_parts = self.n_long = 0
            return
        p64 = ptr.to(torch.int64)
        start, deg = p64[:-1], p64[1:] - p64[:-1]
        is_long = deg > long_segment
        idx = torch.arange(n, device=dev, dtype=torch.int64)
        # a long segment is alone in its chunk: the island id changes at it and right after it
        bump = is_long.clone()
        bump[1:] |= is_long[:-1]
        island = torch.cumsum(bump.to(torch.int64), 0)
        new = torch.ones(n, dtype=torch.bool, device=dev)
        new[1:] = ((island[1:] != island[:-1]) | ((start[1:] // chunk_edges) != (start[:-1] // chunk_edges))
                   | ((idx[1:] // chunk_nodes) != (idx[:-1] // chunk_nodes)))
        heads = torch.nonzero(new).flatten()
        ends = torch.cat([heads[1:], torch.tensor([n], device=dev)])
        head_long = is_long[heads]
        # ordinary chunks
        o_lo, o_nn = heads[~head_long], (ends - heads)[~head_long]
        # split segments -> parts
        ln = heads[head_long]
        n_parts_per = (deg[ln] + part_edges - 1) // part_edges
        self.n_long = int(ln.numel())
        part_ptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(n_parts_per, 0)])
        self.n_parts = int(part_ptr[-1].item())
        owner = torch.repeat_interleave(torch.arange(self.n_long, device=dev), n_parts_per)
        within = torch.arange(self.n_parts, device=dev) - part_ptr[:-1][owner]
        pe_lo = start[ln][owner] + within * part_edges
        pe_hi = torch.minimum(pe_lo + part_edges, p64[1:][ln][owner])
        slots = torch.arange(self.n_parts, device=dev)
        c_lo = torch.cat([o_lo, ln[owner]])
        c_nn = torch.cat([o_nn, torch.ones(self.n_parts, dtype=torch.int64, device=dev)])
        c_part = torch.cat([torch.full_like(o_lo, -1), slots])
        # long parts first: they are the longest tasks of the launch
        order = torch.argsort(c_part >= 0, descending=True, stable=True)
        self.chunks = torch.stack([c_lo, c_nn, c_part, torch.zeros_like(c_lo)], 1)[order].to(torch.int32).contiguous()
        self.parts = torch.stack([pe_lo, pe_hi], 1).to(torch.int32).contiguous() if self.n_parts else \
            torch.zeros((0, 2), **i32)
        self.long_node = ln.to(torch.int32)
        self.long_part_ptr = part_ptr.to(torch.int32)
        self.n_chunks = int(self.chunks.size(0))

    # -- the same tables built by the device-side builder (csrc/stream_chunks.cu): no host read until ``finish`` ----
    @classmethod
    def launch(cls, ptr: torch.Tensor, n_edges: int, chunk_edges: int = STREAM_CHUNK_EDGES,
               chunk_nodes: int = STREAM_CHUNK_NODES, long_segment: int = LONG_SEGMENT,
               part_edges: int = PART_EDGES) -> "StreamChunks":
        """Enqueue the build on the current stream; ``counts`` (int32 [3] on the device) holds (n_chunks, n_parts,
        n_long) once it ran.  Call ``finish(counts_on_host)`` after reading them back."""
        self = cls.__new__(cls)
        dev = ptr.device
        n = ptr.numel() - 1
        i32 = dict(dtype=torch.int32, device=dev)
        max_long = n_edges // (long_segment + 1) + 1
        max_parts = n_edges // part_edges + max_long
        max_chunks = max(n, 0) + max_parts
        self.chunks = torch.empty((max_chunks, 4), **i32)
        self.parts = torch.empty((max_parts, 2), **i32)
        self.long_node = torch.empty((max_long,), **i32)
        self.long_part_ptr = torch.empty((max_long + 1,), **i32)
        self.counts = torch.empty((4,), **i32)
        lib = _lib.load()
        ws_bytes = int(lib.relgat_stream_chunks_workspace_bytes(max(n, 0)))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.relgat_stream_chunks_build(
                _lib.ptr(ptr), max(n, 0), chunk_edges, chunk_nodes, long_segment, part_edges,
                _lib.ptr(self.chunks), max_chunks, _lib.ptr(self.parts), max_parts, _lib.ptr(self.long_node),
                _lib.ptr(self.long_part_ptr), max_long, _lib.ptr(self.counts), _lib.ptr(ws), ws_bytes,
                torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "relgat_stream_chunks_build")
        self.n_chunks = self.n_parts = self.n_long = None
        return self

    @classmethod
    def launch_for_rows(cls, ptr: torch.Tensor, n_edges: int, rows: torch.Tensor, n_rows_dev: torch.Tensor,
                        long_segment: int = LONG_SEGMENT, part_edges: int = PART_EDGES) -> "StreamChunks":
        """Table over the listed segments only (``rows`` int64, ascending; its true length is the device int
        ``n_rows_dev``, ``rows.numel()`` an upper bound): one chunk per listed segment.  ``finish`` as for ``launch``."""
        self = cls.__new__(cls)
        dev = ptr.device
        n = int(rows.numel())
        i32 = dict(dtype=torch.int32, device=dev)
        max_long = n_edges // (long_segment + 1) + 1
        max_parts = n_edges // part_edges + max_long
        max_chunks = n + max_parts
        self.chunks = torch.empty((max_chunks, 4), **i32)
        self.parts = torch.empty((max_parts, 2), **i32)
        self.long_node = torch.empty((max_long,), **i32)
        self.long_part_ptr = torch.empty((max_long + 1,), **i32)
        self.counts = torch.empty((4,), **i32)
        lib = _lib.load()
        ws_bytes = int(lib.relgat_stream_chunks_workspace_bytes(n))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = lib.relgat_stream_chunks_for_rows(
                _lib.ptr(ptr), _lib.ptr(rows), n, _lib.ptr(n_rows_dev), long_segment, part_edges,
                _lib.ptr(self.chunks), max_chunks, _lib.ptr(self.parts), max_parts, _lib.ptr(self.long_node),
                _lib.ptr(self.long_part_ptr), max_long, _lib.ptr(self.counts), _lib.ptr(ws), ws_bytes,
                torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "relgat_stream_chunks_for_rows")
        self.n_chunks = self.n_parts = self.n_long = None
        return self

    def finish(self, counts) -> "StreamChunks":
        n_chunks, n_parts, n_long = (int(v) for v in counts[:3])
        self.n_edges = int(counts[3]) if len(counts) > 3 else None
        if n_chunks < 0:
            raise RuntimeError("relgat_stream_chunks_build: table bounds exceeded")
        self.n_chunks, self.n_parts, self.n_long = n_chunks, n_parts, n_long
        self.chunks = self.chunks[:n_chunks]
        self.parts = self.parts[:n_parts]
        self.long_node = self.long_node[:n_long]
        self.long_part_ptr = self.long_part_ptr[:n_long + 1]
        return self


class GraphIndex:
    """All integer structures of one message-passing graph, as int32 CUDA tensors."""

    def __init__(self, edge_index: torch.Tensor, edge_type: torch.Tensor, num_nodes: int, num_rel: int,
                 validate: bool = True, num_src_nodes: Optional[int] = None,
                 fwd_chunks: bool = True, src_chunks: bool = True, degrees: bool = True, lean: bool = False):
        """``num_nodes`` = destination rows (segments); ``num_src_nodes`` = rows of the feature matrix
        the sources index (defaults to ``num_nodes``; larger on a destination-range partition).
        ``fwd_chunks`` / ``src_chunks``: build the work tables of the forward / by-source kernel (a
        partitioned graph that only ever runs one of the two skips the other).  ``degrees``: also read back the maximum
        in / out degree (two host syncs; the per-batch blocks skip them).  ``lean``: the per-step path — no validation,
        no degrees, work tables from the device-side builder, ONE host read for all the sizes the launches need."""
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
        if validate and not lean and E > 0:
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
        self.max_in_degree = self.max_out_degree = None
        if lean:
            self.fwd_chunks = StreamChunks.launch(self.rowptr, E) if fwd_chunks else None
            self.src_chunks = StreamChunks.launch(self.colptr, E) if src_chunks else None
            pend = [ck for ck in (self.fwd_chunks, self.src_chunks) if ck is not None]
            host = torch.cat([ck.counts for ck in pend] + [self.relptr]).cpu().numpy()  # the one host read
            for k, ck in enumerate(pend):
                ck.finish(host[4 * k:4 * k + 4])
            self._build_rel_chunks(host[4 * len(pend):])
            del ws
            return
        self._build_rel_chunks()
        self.fwd_chunks = StreamChunks(self.rowptr) if fwd_chunks else None
        self.src_chunks = StreamChunks(self.colptr) if src_chunks else None
        if degrees:
            self.max_in_degree = int((self.rowptr[1:] - self.rowptr[:-1]).max().item()) if N > 0 else 0
            self.max_out_degree = int((self.colptr[1:] - self.colptr[:-1]).max().item()) if NS > 0 else 0
        del ws

    def _build_rel_chunks(self, relptr_host=None) -> None:
        """Fixed-size chunks of the by-relation order; a chunk never spans two relations."""
        import numpy as np
        relptr = (self.relptr.cpu().numpy() if relptr_host is None else np.asarray(relptr_host)).astype(np.int64)
        a, b = relptr[:-1], relptr[1:]
        per = (b - a + REL_CHUNK - 1) // REL_CHUNK          # chunks of each relation
        cptr = np.concatenate([[0], np.cumsum(per)])
        n = int(cptr[-1])
        owner = np.repeat(np.arange(self.R), per)
        lo = a[owner] + (np.arange(n) - cptr[:-1][owner]) * REL_CHUNK
        hi = np.minimum(lo + REL_CHUNK, b[owner])
        self.n_chunks = n
        packed = torch.from_numpy(np.concatenate([lo, hi, cptr]).astype(np.int32)).to(self.device)  # one upload
        self.chunk_lo, self.chunk_hi, self.rel_chunk_ptr = packed[:n], packed[n:2 * n], packed[2 * n:]

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
