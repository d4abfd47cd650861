"""Receptive-field blocks of a batch.

The loss of a training step reads the stack's output for the <= B*(2+K) nodes its triples name (reference
model.py:136-137).  Layer L's output for those nodes depends on layer L-1's output for the sources of their in-edges
only, and so on down: the L-hop in-neighbourhood.  ``build_blocks`` extracts, for one batch, one bipartite block per
layer — block l maps the rows D_{l-1} (its sources) to the rows D_l (its destinations, every one with ALL its in-edges in
their full-graph order, so softmax and aggregation see exactly the terms and the order of the full pass) — and the
ordinary kernels run on the blocks.  Loss and parameter gradients equal the full-graph step's; the work is the batch's
receptive field instead of the whole graph, which is why bench.py reports it as a separately labelled record and not
under the headline metric.

The reference has no such path (it always runs all N nodes and E edges, model.py:274-292).
"""
from __future__ import annotations

from typing import List, NamedTuple

import torch

from .graph import GraphIndex


class Blocks(NamedTuple):
    graphs: List[GraphIndex]   # one per layer, first layer first
    input_rows: torch.Tensor   # int64 [graphs[0].N_src]: rows of the node-embedding matrix the first layer reads (sorted)
    out_pos: torch.Tensor      # int64 [len(ids)]: row of graphs[-1]'s output that holds each requested id
    n_edges: int               # edges over all blocks (the work of one pass over the stack)


def _record(g: GraphIndex, stream: torch.cuda.Stream) -> None:
    for v in list(vars(g).values()) + [t for ck in (g.fwd_chunks, g.src_chunks) if ck is not None for t in vars(ck).values()]:
        if isinstance(v, torch.Tensor) and v.is_cuda:
            v.record_stream(stream)


def build_blocks(full: GraphIndex, ids: torch.Tensor, num_layers: int) -> Blocks:
    """Blocks for the rows ``ids`` (int64, may repeat) of an L-layer stack over ``full``.  Runs on the caller's side
    stream: the few host reads (block sizes) then wait for these small kernels only, not for the step before."""
    from .functional import _side_stream
    if ids.dtype != torch.int64 or ids.dim() != 1:
        raise TypeError("ids must be a 1-D int64 tensor")
    if num_layers < 1:
        raise ValueError("num_layers must be >= 1")
    dev = ids.device
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_stream(main)
    ids.record_stream(side)
    with torch.cuda.stream(side):
        D = torch.unique(ids)  # sorted
        if D.numel() == 0:  # nothing requested: one (unread) destination keeps every block non-empty
            D = torch.zeros(1, dtype=torch.int64, device=dev)
        out_pos = torch.searchsorted(D, ids)
        graphs: List[GraphIndex] = []
        n_edges = 0
        rowptr = full.rowptr
        for _ in range(num_layers):
            lo = rowptr[D].to(torch.int64)
            cnt = rowptr[D + 1].to(torch.int64) - lo
            n_e = int(cnt.sum().item())
            dst_local = torch.repeat_interleave(torch.arange(D.numel(), device=dev), cnt, output_size=n_e)
            first = torch.cumsum(cnt, 0) - cnt
            pos = torch.arange(n_e, device=dev) - first[dst_local] + lo[dst_local]
            src_global = full.csr_src[pos].to(torch.int64)
            rel = full.csr_rel[pos].to(torch.int64)
            S = torch.unique(src_global)
            if S.numel() == 0:  # destinations without in-edges: keep one (unread) source row so no layer has 0 rows
                S = torch.zeros(1, dtype=torch.int64, device=dev)
            src_local = torch.searchsorted(S, src_global)
            g = GraphIndex(torch.stack([src_local, dst_local]), rel, int(D.numel()), full.R, validate=False,
                           num_src_nodes=int(S.numel()), lean=True)
            graphs.append(g)
            n_edges += n_e
            D = S
        graphs.reverse()
    main.wait_stream(side)
    for g in graphs:
        _record(g, main)
    for t in (D, out_pos):
        t.record_stream(main)
    return Blocks(graphs, D, out_pos, n_edges)
