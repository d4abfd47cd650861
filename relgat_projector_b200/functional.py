"""Autograd wiring of the RelGAT hot path over the sm_100a kernels.

``RelGATStackFunction`` runs L RelGAT layers (with the inter-layer ELU of reference
core/model/model.py:286-287 fused into the edge kernel's epilogue) as ONE autograd node:
forward = [split W -> tcgen05 GEMM -> fused edge kernel] per layer, backward = the closed
form of SURVEY.md §A.2 with the by-source / by-relation passes and the tcgen05 dW / dX GEMMs.
Nothing here does arithmetic in torch; torch allocates and orders the launches.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch

from . import ops
from .graph import GraphIndex

PRECISIONS = ("fp32", "bf16")
# dA through the logit-table gradient (dS columns beside dP, widened dW GEMM) instead of the by-relation gather pass
# (SURVEY.md A.3); RELGAT_DS=0 restores the by-relation kernel (kept for the partitioned path and for A/B timing)
USE_DS = os.environ.get("RELGAT_DS", "1") != "0"
# backward prep of a hidden layer fused into the epilogue of the dX GEMM above it (RELGAT_FUSE_PREP=0: separate kernel)
# OFF by default: measured on config 2, the fused epilogue costs 3.3 ms against 1.30 (GEMM) + 0.61 (prep kernel) — four
# epilogue warps reading y row-wise cannot keep up with the MMA (profiles/r02_summary.md); kept as an experiment knob.
FUSE_PREP = os.environ.get("RELGAT_FUSE_PREP", "0") != "0"
# the small tail of the dS path (dA = T W^T, dbeta: ~60 us per layer) in line on the main stream (0, default) or on the
# side stream (1); measured alike (13.48 vs 13.52 ms per step), in line keeps the per-kernel timings of bench.py exact
TAIL_ON_SIDE = os.environ.get("RELGAT_TAIL_SIDE", "0") != "0"
# exact-zero rows of the output gradient: the loss reads <= B*(2+K) rows of the stack's output, so dL/d out_L is zero
# outside them and dL/d out_l is zero outside the sources of the edges into layer l+1's non-zero rows.  The by-source
# pass skips the edges into rows known to be zero (their dz and their contribution to dP are exact zeros): same gradient,
# fewer gathers.  RELGAT_SPARSE_BWD=0 gathers every edge (bench.py's headline does, so that its unit of work stays the
# dense pass of SURVEY.md §8(d); the sparse pass is reported beside it).
SPARSE_BWD = os.environ.get("RELGAT_SPARSE_BWD", "1") != "0"
# ... and, on top of it (fp32 mode), the backward is COMPACTED to those rows: the by-source pass writes only the dP rows
# of sources with an edge into a non-zero row, the weight-gradient / dX GEMMs run over those rows only, the hidden layers'
# prep touches only them.  The row sets depend on the batch ids and the graph alone, so they (and their sizes, which the
# host needs for the GEMM shapes) are prepared on the side stream during the forward.  RELGAT_COMPACT_BWD=0: skip edges
# only (dense rows).
COMPACT_BWD = os.environ.get("RELGAT_COMPACT_BWD", "1") != "0"

# third-generation by-source pass (csrc/edge_bwd_src3.cu, fp32 mode): gathered rows as bulk async copies into a per-warp
# shared-memory ring, rows [dPa | dS] WITHOUT the per-edge dz * A[rel] term; dP = dPa + dS·A is folded into the GEMMs
# that consume the rows (fold_operands below).  1.41 -> 1.12 ms per launch at config 2.  RELGAT_SRC_V3=0: first generation.
SRC_V3 = os.environ.get("RELGAT_SRC_V3", "1") != "0"

# SM split of the backward (fp32 mode, third-generation by-source pass): the weight-gradient GEMM of layer l (tensor-
# bound, little HBM traffic) runs on OVERLAP_SMS SMs of the side stream while the HBM-bound prep + by-source pass of layer
# l-1 run on the remaining SMs of the main stream (they depend on dX, not on dW).  0 = everything in line.
OVERLAP_SMS = int(os.environ.get("RELGAT_OVERLAP_SMS", "0"))

_SIDE_STREAMS = {}


def _side_stream(device) -> torch.cuda.Stream:
    """One extra stream per device: the HBM-bound by-relation pass runs beside the tensor-bound
    dW / dX GEMMs (they only share read-only inputs)."""
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device, priority=-1)  # its CTAs are placed first when SMs free up
    return _SIDE_STREAMS[key]


# ---------------------------------------------------------------------------------------------
# sparse-gradient hint: the scorer's backward produces a dense [N, D] gradient whose non-zero rows are the
# <= B*(2+K) batch rows (reference model.py:136-137: x[src_ids], x[dst_ids]).  It tags that tensor with the
# row list; the stack's backward uses the tag to compute t / hsum of the last layer from those rows only
# (every other row is an exact zero either way).  The tag is ignored unless the tensor is untouched since
# it was tagged (autograd may accumulate other gradients into it in place: the version counter tells).
# ---------------------------------------------------------------------------------------------
def mark_sparse_rows(grad: torch.Tensor, rows: torch.Tensor) -> None:
    grad._relgat_rows = (rows, grad._version)


def sparse_rows_of(grad: torch.Tensor) -> Optional[torch.Tensor]:
    tag = getattr(grad, "_relgat_rows", None)
    if tag is None or tag[1] != grad._version:
        return None
    return tag[0]


# ---------------------------------------------------------------------------------------------
# persistent all-zero [N, C] tables.  The loss reads <= B*(2+K) rows of the stack's output (reference model.py:136-137),
# so the gradient of that output is zero outside those rows.  The fused "stack + row gather" node scatters the batch
# rows' gradients into such a table, runs the last layer's backward on it (the by-source pass gathers G[dst] for every
# edge, so a dense table is needed) and clears the rows again: no [N, C] memset and no dense index_add per step.
# ---------------------------------------------------------------------------------------------
_ZERO_TABLES = {}


def _take_zero_table(device, n: int, c: int) -> torch.Tensor:
    key = (torch.device(device).index, int(n), int(c))
    pool = _ZERO_TABLES.setdefault(key, [])
    return pool.pop() if pool else torch.zeros((n, c), dtype=torch.float32, device=device)


def _return_zero_table(t: torch.Tensor) -> None:
    key = (t.device.index, int(t.size(0)), int(t.size(1)))
    pool = _ZERO_TABLES.setdefault(key, [])
    if len(pool) < 2:  # all rows are zero again (stream-ordered): ready for the next backward on this stream
        pool.append(t)


def clear_zero_tables() -> None:
    _ZERO_TABLES.clear()


def stable_sort_ids(ids: torch.Tensor, n_max: int):
    """(sorted keys int64, perm int64) of a stable sort; the radix sort runs on the narrowest integer type that holds
    ``n_max`` (CUB does 8 passes over int64 keys, 2 over int16: ~100 us -> ~35 us for a batch's ids)."""
    narrow = torch.int16 if n_max < 2 ** 15 else (torch.int32 if n_max < 2 ** 31 else torch.int64)
    keys, perm = torch.sort(ids.to(narrow) if narrow != ids.dtype else ids, stable=True)
    return keys.to(torch.int64), perm


def presort_on_side_stream(ids: torch.Tensor, n_max: int):
    """Stable sort of a batch's id vector on the side stream (it is only needed by the backward: off the forward's
    critical chain).  Returns (keys, perm, event); wait for ``event`` on the consuming stream before use."""
    main = torch.cuda.current_stream(ids.device)
    side = _side_stream(ids.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        keys, perm = stable_sort_ids(ids, n_max)
        ev = torch.cuda.Event()
        ev.record(side)
    for t in (keys, perm):
        t.record_stream(main)
    ids.record_stream(side)
    return keys, perm, ev


def fold_operands(A: torch.Tensor, Wp, WTp, H: int, F: int, R: int, d_in: int):
    """GEMM operands that fold dP = dPa + dS·A_bd (A_bd [H*R, C]: block-diagonal attention vectors, row h*R + r holds
    A[h, r, :] in columns h*F .. (h+1)*F) into the consumers of the rows [dPa | dS] the third-generation by-source pass
    writes — parameters only, prepared once per step and layer in the forward:
      * ``Abd``: planes of A_bd;  dW = dPaᵀX + A_bdᵀ (dSᵀX)  (weight_grad_gemm gives both products in one launch);
      * ``Bext`` (layers with a dX): planes of [Wᵀ | (A_bd·W)ᵀ | 0] [d_in, Wd]:  dX = [dPa | dS] · [W ; A_bd·W]."""
    C, HR = H * F, H * R
    Wd = ops.ds_row_width(H, F, R)
    Abd = torch.zeros((HR, C), dtype=torch.float32, device=A.device)
    idx = torch.arange(H, device=A.device)
    Abd.view(H, R, H, F)[idx, :, idx, :] = A
    Abd_p = ops.split_bf16(Abd, True)
    Bext = None
    if WTp is not None:
        AW = ops.gemm(Abd_p, False, Wp, True, HR, d_in, C)  # [H*R, d_in] = A_bd · W
        AWt_p = ops.split_bf16(AW.t().contiguous(), True)
        pad = [WTp[0].new_zeros((d_in, Wd - C - HR))] if Wd > C + HR else []
        Bext = tuple(torch.cat([w_, a_] + pad, dim=1) for w_, a_ in zip(WTp, AWt_p))
    return dict(Abd=Abd_p, Bext=Bext)


def weight_grad_gemm(dPp, xp, Wd: int, d_in: int, rows: int, device) -> torch.Tensor:
    """dW_ext [Wd, d_in] = [dP | dS]^T · X over ``rows`` rows (split-K, both operands MN-major).  Which operand supplies
    the M dimension is chosen by the padded-tile cost model: at config 2's second layer (Wd = 1000, d_in = 800) the
    transposed product runs 256-wide N tiles instead of 160-wide ones (1.36 vs 1.61 ms)."""
    if ops.gemm_cost_model(d_in, Wd) < 0.97 * ops.gemm_cost_model(Wd, d_in):
        t = ops.gemm(xp, True, dPp, True, d_in, Wd, rows, splits_k=ops.pick_splits_k(d_in, Wd, rows, device))
        return t.t().contiguous()
    return ops.gemm(dPp, True, xp, True, Wd, d_in, rows, splits_k=ops.pick_splits_k(Wd, d_in, rows, device))


def _plan_compact_backward(gather, g: GraphIndex, L: int, forward: bool = False):
    """Row sets of the compacted backward, one entry per layer (first layer first): ``dst_bits`` = rows of dL/d out_l
    that can be non-zero, ``rank`` / ``list`` = compact numbering of the sources of the edges into them (= the rows of
    dP_l, and the non-zero rows one layer down), ``counts`` = their sizes in pinned host memory once ``ready`` fired.
    ``forward``: also ``fwd_chunks`` = the forward kernel's work table over exactly the destinations layer l must produce
    (the receptive-field forward).  Everything runs on the side stream behind the id sort: off the critical chain."""
    from .graph import StreamChunks
    keys, _, sorted_ev = gather
    dev = keys.device
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev)
    side.wait_event(sorted_ev)
    with torch.cuda.stream(side):
        bits = ops.mark_rows(keys, g.N)
        dst_list = dst_count = None
        if forward:
            _, dst_list, dst_count = ops.bitmap_ranks(bits, g.N)
        layers = []
        for _ in range(L):
            sbits = ops.mark_sources(bits, g)
            rank, lst, cnt = ops.bitmap_ranks(sbits, g.N_src)
            p = dict(dst_bits=bits, rank=rank, list=lst, count=cnt)
            if forward:
                p["fwd_chunks"] = StreamChunks.launch_for_rows(g.rowptr, g.E, dst_list, dst_count)
                dst_list, dst_count = lst, cnt
            layers.append(p)
            bits = sbits
        layers.reverse()
        dev_counts = [p["count"] for p in layers] + ([p["fwd_chunks"].counts for p in layers] if forward else [])
        counts = torch.empty((L * (5 if forward else 1),), dtype=torch.int32, pin_memory=True)
        counts.copy_(torch.cat(dev_counts), non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(side)
    for p in layers:
        for k in ("dst_bits", "rank", "list"):
            p[k].record_stream(main)
        if forward:
            for t in vars(p["fwd_chunks"]).values():
                if isinstance(t, torch.Tensor):
                    t.record_stream(main)
    return dict(layers=layers, counts=counts, ready=ready, n_layers=L)


LAST_PRUNED_EDGES = None  # layer-edges processed by the latest receptive-field forward (diagnostic, bench.py)


def _plan_counts(plan) -> List[int]:
    """Host copy of the plan's sizes (waits for the side stream's few small kernels, not for the main stream) and,
    for a forward plan, the work tables cut to their true sizes."""
    if "sizes" not in plan:
        plan["ready"].synchronize()
        host = [int(v) for v in plan["counts"].tolist()]
        L = plan["n_layers"]
        plan["sizes"] = host[:L]
        for l, p in enumerate(plan["layers"]):
            if "fwd_chunks" in p:
                p["fwd_chunks"].finish(host[L + 4 * l:L + 4 * l + 4])
        if plan["layers"] and "fwd_chunks" in plan["layers"][0]:
            global LAST_PRUNED_EDGES
            LAST_PRUNED_EDGES = sum(p["fwd_chunks"].n_edges for p in plan["layers"])
    return plan["sizes"]


class LayerDropout:
    """Dropout state of one layer for one forward/backward: ``feat`` = ops.DropMask of the feature dropout on the
    layer's output (reference layer.py:321-322) or None, ``edge`` = ops.DropMask of the attention dropout
    (layer.py:296-297) or None."""

    def __init__(self, feat=None, edge=None):
        self.feat, self.edge = feat, edge


class RelGATStackFunction(torch.autograd.Function):
    """out = RelGAT_L(... ELU(RelGAT_1(x0)) ...) for layers sharing one graph.

    args: x0 [N, D_in] fp32, then per layer (W [H*F, D_in_l], A [H, R, F], beta [R] or None).
    ``x0_planes`` may carry a cached bf16 split of a frozen x0 (reference model.py:32 keeps the
    input embeddings as a buffer, so the split is paid once, not per step).
    ``drop``: None or one LayerDropout per layer (masks applied inside the edge kernels).
    ``gather_ids``: None = return all N rows; int64 [n] = return out[gather_ids] only (the rows the scorer reads,
    reference model.py:136-137) — backward then receives [n, C] and never builds a dense [N, C] gradient.
    ``graph``: one GraphIndex shared by the layers, or a list with one bipartite block per layer (blocks.py: layer l
    maps block.N_src input rows to block.N output rows, and block l+1's sources are block l's destinations).
    """

    @staticmethod
    def forward(ctx, x0, graph: GraphIndex, heads: int, out_dim: int, precision: str, x0_planes, drop, gather_ids,
                prune, *params):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        if len(params) % 3 != 0 or not params:
            raise ValueError("params must be (W, A, beta) per layer")
        L = len(params) // 3
        H, F = heads, out_dim
        C = H * F
        blocks = isinstance(graph, (list, tuple))
        graphs = list(graph) if blocks else [graph] * L
        if len(graphs) != L:
            raise ValueError(f"{len(graphs)} blocks for {L} layers")
        for l in range(L):
            want = graphs[l - 1].N if l else (x0_planes[0].size(0) if x0 is None else x0.size(0))
            if graphs[l].N_src != want:
                raise ValueError(f"layer {l}: {want} input rows but the graph's sources index {graphs[l].N_src} rows")
        with_lo = precision == "fp32"
        planes = x0_planes if x0_planes is not None else ops.split_bf16(x0, with_lo)
        saved = []
        out = None
        ctx.gather = ctx.plan = None
        # receptive-field forward: only the rows the batch's rows depend on are produced (see _plan_compact_backward)
        prune = bool(prune) and gather_ids is not None and USE_DS and with_lo and not blocks and \
            not (x0 is not None and x0.requires_grad)
        sizes = None
        if prune:
            ctx.gather = presort_on_side_stream(gather_ids.contiguous(), graphs[-1].N)
            ctx.plan = _plan_compact_backward(ctx.gather, graphs[-1], L, forward=True)
            sizes = _plan_counts(ctx.plan)
            torch.cuda.current_stream(planes[0].device).wait_event(ctx.plan["ready"])
        ctx.pruned = prune
        if gather_ids is not None and ctx.gather is None:
            # (sorted keys, perm, event) = the summation order of backward: sorted on the side stream beside the forward
            ctx.gather = presort_on_side_stream(gather_ids.contiguous(), graphs[-1].N)
        for l in range(L):
            W, A, beta = params[3 * l], params[3 * l + 1], params[3 * l + 2]
            d_in = W.size(1)
            if W.size(0) != C:
                raise ValueError(f"layer {l}: W must be [{C}, D_in]")
            gl = graphs[l]
            x0_grad = x0 is not None and x0.requires_grad
            Wp = ops.split_bf16(W.detach(), with_lo)
            # K-major copy of Wᵀ for dX = dP·W (3 MB transpose; the K-major B path is ~12% faster than MN-major)
            WTp = ops.split_bf16(W.detach().t().contiguous(), with_lo) if (l > 0 or x0_grad) else None
            fold = None
            if SRC_V3 and USE_DS and with_lo and F % 4 == 0:
                # parameters only: prepared on the side stream beside this layer's GEMM, waited for in backward
                main, side = torch.cuda.current_stream(W.device), _side_stream(W.device)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    fold = fold_operands(A.detach(), Wp, WTp, H, F, gl.R, d_in)
                    fold["ready"] = torch.cuda.Event()
                    fold["ready"].record(side)
                for pp in (Wp, WTp, fold["Abd"], fold["Bext"]):
                    for tns in (pp or ()):
                        if tns is not None:
                            tns.record_stream(side)
                            tns.record_stream(main)
            # "bf16": projected features are stored in bf16 (halves every gather of the edge kernels)
            last = l == L - 1
            dl = drop[l] if drop is not None else None
            pl = ctx.plan["layers"][l] if prune else None
            if prune:
                # the input rows this layer's destinations read -> compact planes -> compact P (row rank[src])
                n_s = sizes[l]
                if n_s > 0:
                    planes = tuple(None if p_ is None else ops.gather_plane_rows(p_, pl["list"][:n_s]) for p_ in planes)
                else:  # no edge reaches the batch's rows at this layer: one unread zero row keeps the shapes valid
                    planes = tuple(None if p_ is None else p_.new_zeros((1, d_in)) for p_ in planes)
            P = ops.gemm(planes, False, Wp, False, planes[0].size(0) if prune else gl.N_src, C, d_in,
                         out_dtype=torch.float32 if with_lo else torch.bfloat16)
            out, act, _, z, minv, bias = ops.edge_fwd(P, A.detach(), None if beta is None else beta.detach(), gl,
                                                      H, F, want_act=not last, apply_elu=True, act_lo=with_lo,
                                                      feat_drop=dl.feat if dl else None, edge_drop=dl.edge if dl else None,
                                                      chunks=pl["fwd_chunks"] if prune else None,
                                                      src_row=pl["rank"] if prune else None)
            saved.append(dict(xp=planes, Wp=Wp, WTp=WTp, P=P, out=out, minv=minv, z=z, bias=bias, A=A.detach(),
                              d_in=d_in, has_beta=beta is not None, drop=dl, fold=fold))
            planes = act
        ctx.saved = saved
        ctx.graphs = graphs
        ctx.blocks = blocks
        ctx.cfg = (H, F, L, with_lo)
        ctx.x0_needs_grad = bool(x0 is not None and x0.requires_grad)
        if gather_ids is not None:
            ids = gather_ids.contiguous()
            if ctx.gather is None:
                ctx.gather = presort_on_side_stream(ids, graphs[-1].N)  # (sorted keys, perm, event): summation order of backward
            if ctx.plan is None and SPARSE_BWD and COMPACT_BWD and USE_DS and with_lo and not blocks and \
                    not ctx.x0_needs_grad:
                ctx.plan = _plan_compact_backward(ctx.gather, graphs[-1], L)
            rows = out.new_empty((ids.numel(), C))
            ops.pull_rows(out, ids, rows)
            return rows
        return out

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.saved is None:
            raise RuntimeError("the RelGAT stack's saved state was released by a previous backward; re-run the forward "
                               "(retain_graph is not supported by the fused stack)")
        H, F, L, with_lo = ctx.cfg
        graphs, blocks = ctx.graphs, ctx.blocks
        C = H * F
        N = graphs[-1].N  # rows of the stack's output
        grads: List[Optional[torch.Tensor]] = [None] * (3 * L)
        table = None
        if ctx.gather is not None:
            # gradient of the gathered rows -> rows of a persistent zero table (ordered segmented sum: a node may be
            # named several times by the batch); everything else stays exactly zero
            keys, perm, ready = ctx.gather
            torch.cuda.current_stream(grad_out.device).wait_event(ready)
            # blocks: the last block has only the batch's own rows (a different count every step): a fresh small table
            table = (torch.zeros((N, C), dtype=torch.float32, device=grad_out.device) if blocks
                     else _take_zero_table(grad_out.device, N, C))
            ops.index_add_sorted(grad_out.contiguous(), keys, N, out=table, presorted=(keys, perm), accumulate=False)
            dY, nz_rows, owned = table, (None if blocks else keys), True
        else:
            dY = grad_out.contiguous()
            nz_rows = sparse_rows_of(grad_out) if dY is grad_out else None
            owned = False
        if ctx.plan is not None and table is not None and (ctx.pruned or (SPARSE_BWD and COMPACT_BWD)):
            grads = _backward_compact(ctx, table, keys)
            ctx.saved = None
            return (None, None, None, None, None, None, None, None, None, *grads)
        # (blocks hold nothing but the rows the batch reaches: there is nothing to skip)
        nz_bits = ops.mark_rows(nz_rows, N) if (SPARSE_BWD and USE_DS and nz_rows is not None and not blocks) else None
        dX = None
        main_sms = None  # SM budget of the main stream while the layer above's dW GEMM runs beside it (OVERLAP_SMS)
        prepped = None  # (G, t, hsum) of layer l when the dX GEMM of layer l+1 produced them in its epilogue
        fuse_prep = USE_DS and FUSE_PREP and with_lo and ops.gemm_dx_prep_supported(C, F)
        for l in reversed(range(L)):
            s = ctx.saved[l]
            g = graphs[l]
            n_src = g.N_src
            dl = s["drop"]
            with ops.sm_limit(main_sms):
                if prepped is not None:
                    G, t, hsum = prepped
                    prepped = None
                else:
                    # fp32 storage: G aliases dY and only the batch rows are touched; bf16 storage writes a dense bf16 G
                    G, t, hsum = ops.edge_bwd_prep(dY, s["out"], s["bias"], H, F, apply_elu=(l < L - 1), inplace=owned,
                                                   g_bf16=not with_lo, nonzero_rows=nz_rows if l == L - 1 else None,
                                                   feat_drop=dl.feat if dl else None)
                fold = s["fold"] if (USE_DS and ops.src3_supported(G, F) and ops.src3_supported(s["P"], F)) else None
                if fold is not None:  # rows [dPa | dS], dS·A folded into the GEMMs below
                    torch.cuda.current_stream(dY.device).wait_event(fold["ready"])
                _, dPp, dz = ops.edge_bwd_src(s["P"], G, s["A"], s["z"], s["minv"], t, g, H, F,
                                              want_fp32=False, want_planes=True, planes_lo=with_lo,
                                              edge_drop=dl.edge if dl else None, want_ds=USE_DS, dst_nz=nz_bits,
                                              a_term=fold is None)
            if main_sms is not None:
                torch.cuda.current_stream(dY.device).wait_stream(_side_stream(dY.device))  # the GEMMs below want the whole chip
                main_sms = None
            if nz_bits is not None and l > 0:
                nz_bits = ops.mark_sources(nz_bits, g)  # rows of dP, hence of dL/d out_{l-1}, that can be non-zero
            if table is not None and l == L - 1:
                if not blocks:
                    ops.zero_rows(table, nz_rows)  # the table's rows are consumed (G aliased it): all-zero again
                    _return_zero_table(table)
                G = None
            d_in = s["d_in"]
            main = torch.cuda.current_stream(dY.device)
            side = _side_stream(dY.device)
            if USE_DS:
                # widened rows [dP | dS]: ONE split-K GEMM gives dW (first C rows) and dS^T X (the H*R rows below);
                # dA[h] = (dS_h^T X) W_h^T is a 200 x 800 x d_in GEMM; dbeta needs hsum only.  No gather of P.
                HR = H * g.R
                Wd = dPp[0].size(1)
                dP_c = tuple(None if p_ is None else p_[:, :C] for p_ in dPp)
                if OVERLAP_SMS > 0 and l > 0 and fold is not None and not TAIL_ON_SIDE:
                    # dX first (whole chip), then dW + its tail on OVERLAP_SMS SMs of the side stream beside the layer
                    # below's prep and by-source pass on the others
                    dX = ops.gemm(dPp, False, fold["Bext"], False, n_src, d_in, Wd)
                    dx_done = torch.cuda.Event()
                    dx_done.record(main)
                    side.wait_event(dx_done)
                    total = ops.sm_count(dY.device)
                    k_side = max(2, min(total - 2, OVERLAP_SMS)) // 2 * 2
                    with torch.cuda.stream(side), ops.sm_limit(k_side):
                        dW_ext = weight_grad_gemm(dPp, s["xp"], Wd, d_in, n_src, dY.device)
                        Tp = ops.split_bf16(dW_ext[C:C + HR].contiguous(), with_lo)
                        dW = dW_ext[:C] + ops.gemm(fold["Abd"], True, Tp, True, C, d_in, HR)
                        dA_full = ops.gemm(Tp, False, s["Wp"], False, HR, C, d_in)
                        dA = torch.stack([dA_full[h * g.R:(h + 1) * g.R, h * F:(h + 1) * F] for h in range(H)])
                        dbeta = ops.edge_bwd_beta(hsum, g, H) if s["has_beta"] else None
                    for tns in (*dPp, hsum):
                        if tns is not None:
                            tns.record_stream(side)
                    for tns in (dW, dA, dbeta):
                        if tns is not None:
                            tns.record_stream(main)
                    main_sms = total - k_side
                    grads[3 * l], grads[3 * l + 1], grads[3 * l + 2] = dW, dA, dbeta
                    dY, owned = dX, True
                    del G, dPp, dz
                    continue
                dW_ext = weight_grad_gemm(dPp, s["xp"], Wd, d_in, n_src, dY.device)
                dw_ready = torch.cuda.Event()
                dw_ready.record(main)
                dW = dW_ext[:C]
                Tp = None
                if fold is not None:
                    Tp = ops.split_bf16(dW_ext[C:C + HR].contiguous(), with_lo)
                    dW = dW + ops.gemm(fold["Abd"], True, Tp, True, C, d_in, HR)  # + A_bd^T (dS^T X)
                if l > 0 and fuse_prep and fold is None:
                    # dX never reaches memory: the GEMM's epilogue applies ELU'(y), the dropout mask and the row sums
                    # of the layer below (what edge_bwd_prep would do in a second pass over dX and y)
                    below = ctx.saved[l - 1]
                    bd = below["drop"]
                    prepped = ops.gemm_dx_prep(dP_c, s["WTp"], n_src, d_in, C, below["out"], below["bias"], H, F,
                                               apply_elu=True, feat_drop=bd.feat if bd else None)
                    dX = prepped[0]
                elif l > 0 or ctx.x0_needs_grad:
                    if fold is not None:  # dX = [dPa | dS] · [W ; A_bd·W]
                        dX = ops.gemm(dPp, False, fold["Bext"], False, n_src, d_in, Wd)
                    else:
                        dX = ops.gemm(dP_c, False, s["WTp"], False, n_src, d_in, C)

                def tail(Tp=Tp):
                    if Tp is None:
                        Tp = ops.split_bf16(dW_ext[C:C + HR].contiguous(), with_lo)
                    dA_full = ops.gemm(Tp, False, s["Wp"], False, HR, C, d_in)
                    dA_ = torch.stack([dA_full[h * g.R:(h + 1) * g.R, h * F:(h + 1) * F] for h in range(H)])
                    return dA_, (ops.edge_bwd_beta(hsum, g, H) if s["has_beta"] else None)

                if TAIL_ON_SIDE:
                    side.wait_event(dw_ready)
                    with torch.cuda.stream(side):  # small tail work beside the next layer's kernels
                        dA, dbeta = tail()
                    for tns in (dW_ext, hsum):
                        tns.record_stream(side)
                    for tns in (dA, dbeta):  # allocated on the side stream, consumed by the optimizer on the main one
                        if tns is not None:
                            tns.record_stream(main)
                else:
                    dA, dbeta = tail()
                grads[3 * l], grads[3 * l + 1], grads[3 * l + 2] = dW, dA, dbeta
                if l > 0 or ctx.x0_needs_grad:
                    dY, owned = dX, True
                if l == 0 and TAIL_ON_SIDE:
                    main.wait_stream(side)
                del G, dPp, dz
                continue
            side.wait_stream(main)
            with torch.cuda.stream(side):  # dA / dbeta: gather-bound, overlaps the GEMMs below
                dA, dbeta = ops.edge_bwd_rel(s["P"], dz, hsum, g, H, F, want_dbeta=s["has_beta"])
            splits = ops.pick_splits_k(C, d_in, n_src, dY.device)
            dW = ops.gemm(dPp, True, s["xp"], True, C, d_in, n_src, splits_k=splits)
            grads[3 * l], grads[3 * l + 1], grads[3 * l + 2] = dW, dA, dbeta
            if l > 0 or ctx.x0_needs_grad:
                dX = ops.gemm(dPp, False, s["WTp"], False, n_src, d_in, C)
                dY, owned = dX, True
            main.wait_stream(side)  # join before any buffer of this layer is released or reused
            for tns in (dA, dbeta):
                if tns is not None:
                    tns.record_stream(main)
            del G, dPp, dz
        ctx.saved = None
        return (dX if ctx.x0_needs_grad else None, None, None, None, None, None, None, None, None, *grads)


def _backward_compact(ctx, table: torch.Tensor, keys: torch.Tensor) -> List[Optional[torch.Tensor]]:
    """Backward of the stack restricted to the rows that can be non-zero (see COMPACT_BWD): per layer
    sparse prep -> by-source pass over the marked edges, dP rows written compactly -> gather of the matching input rows
    -> [dP | dS]^T X and dP W over those rows only.  ``table`` = the zero table holding the batch rows' gradients."""
    H, F, L, with_lo = ctx.cfg
    g = ctx.graphs[-1]
    C, N, HR = H * F, g.N, H * g.R
    dev = table.device
    plan = ctx.plan
    counts = _plan_counts(plan)  # sizes of the row sets (computed during the forward: long done)
    torch.cuda.current_stream(dev).wait_event(plan["ready"])
    pruned = ctx.pruned  # receptive-field forward: P and the input planes are already compact
    grads: List[Optional[torch.Tensor]] = [None] * (3 * L)
    dX_c, G_table, prev_rows = None, None, None
    for l in reversed(range(L)):
        s, pl, n_s = ctx.saved[l], plan["layers"][l], counts[l]
        dl = s["drop"]
        if l == L - 1:
            G, t, hsum = ops.edge_bwd_prep(table, s["out"], s["bias"], H, F, apply_elu=False, inplace=True,
                                           nonzero_rows=keys, feat_drop=dl.feat if dl else None)
            clear_rows = keys
        else:
            G_table = _take_zero_table(dev, N, C)
            G, t, hsum = ops.edge_bwd_prep(dX_c, s["out"], s["bias"], H, F, apply_elu=True, G_out=G_table,
                                           compact_rows=prev_rows, feat_drop=dl.feat if dl else None)
            clear_rows = prev_rows
        dPp = None
        fold = s["fold"] if (n_s > 0 and ops.src3_supported(G, F) and ops.src3_supported(s["P"], F)) else None
        if fold is not None:
            torch.cuda.current_stream(dev).wait_event(fold["ready"])
        if n_s > 0:
            _, dPp, _ = ops.edge_bwd_src(s["P"], G, s["A"], s["z"], s["minv"], t, g, H, F, want_fp32=False,
                                         want_planes=True, planes_lo=True, edge_drop=dl.edge if dl else None,
                                         want_ds=True, dst_nz=pl["dst_bits"], src_rows=(pl["rank"], n_s),
                                         p_compact=pruned, a_term=fold is None)
        ops.zero_rows(G, clear_rows)  # the table's rows are consumed: all-zero again for the next step
        _return_zero_table(G)
        if n_s == 0:  # no edge reaches a non-zero row: this layer's and every lower layer's gradients are zero
            for k in range(l + 1):
                sk = ctx.saved[k]
                grads[3 * k] = torch.zeros((C, sk["d_in"]), dtype=torch.float32, device=dev)
                grads[3 * k + 1] = torch.zeros_like(sk["A"])
                grads[3 * k + 2] = torch.zeros((g.R,), dtype=torch.float32, device=dev) if sk["has_beta"] else None
            break
        rows = pl["list"][:n_s]
        xp_c = s["xp"] if pruned else tuple(None if p_ is None else ops.gather_plane_rows(p_, rows) for p_ in s["xp"])
        d_in = s["d_in"]
        dW_ext = weight_grad_gemm(dPp, xp_c, dPp[0].size(1), d_in, n_s, dev)
        Tp = ops.split_bf16(dW_ext[C:C + HR].contiguous(), with_lo)
        dW = dW_ext[:C]
        if fold is not None:  # rows are [dPa | dS]: dW += A_bd^T (dS^T X), dX = [dPa | dS] · [W ; A_bd·W]
            dW = dW + ops.gemm(fold["Abd"], True, Tp, True, C, d_in, HR)
        if l > 0:
            if fold is not None:
                dX_c = ops.gemm(dPp, False, fold["Bext"], False, n_s, d_in, dPp[0].size(1))
            else:
                dX_c = ops.gemm(tuple(p_[:, :C] for p_ in dPp), False, s["WTp"], False, n_s, d_in, C)
            prev_rows = rows
        dA_full = ops.gemm(Tp, False, s["Wp"], False, HR, C, d_in)
        grads[3 * l] = dW
        grads[3 * l + 1] = torch.stack([dA_full[h * g.R:(h + 1) * g.R, h * F:(h + 1) * F] for h in range(H)])
        grads[3 * l + 2] = ops.edge_bwd_beta(hsum, g, H) if s["has_beta"] else None
    return grads


def relgat_stack(x0, graph, heads, out_dim, layer_params: Sequence, precision="fp32", x0_planes=None, drop=None,
                 gather_ids=None, prune=False):
    """``prune`` (with ``gather_ids``, fp32 mode): receptive-field step on the full graph's index — per layer only the
    rows the requested rows depend on are produced, and the backward covers the same rows."""
    flat = []
    for W, A, beta in layer_params:
        flat += [W, A, beta]
    return RelGATStackFunction.apply(x0, graph, heads, out_dim, precision, x0_planes, drop, gather_ids, prune, *flat)


class GatherRowsFunction(torch.autograd.Function):
    """rows = x[ids] with a deterministic backward (ordered segmented sum instead of index_add's atomics); the dense
    gradient it returns carries the row list (see mark_sparse_rows)."""

    @staticmethod
    def forward(ctx, x, ids):
        out = x.new_empty((ids.numel(), x.size(1)))
        ops.pull_rows(x.detach().contiguous(), ids, out)
        ctx.save_for_backward(ids)
        ctx.n_rows = x.size(0)
        return out

    @staticmethod
    def backward(ctx, grad_rows):
        (ids,) = ctx.saved_tensors
        dx, keys = ops.index_add_sorted(grad_rows.contiguous(), ids, ctx.n_rows, return_keys=True)
        mark_sparse_rows(dx, keys)
        return dx, None


class GatherScoreFunction(torch.autograd.Function):
    """scores = scorer(x[src_ids], rel_ids, x[dst_ids]) without materialising the gathers
    (reference model.py:136-141); optional side outputs: transform rows for the first
    ``n_transform`` triples (scorer.transform) and the gathered destination rows."""

    @staticmethod
    def forward(ctx, kind, normalize, x, src_ids, dst_ids, rel_emb, rel_ids, n_transform, want_dst_vec):
        ctx.set_materialize_grads(False)
        score, tr, _, dv = ops.score_fwd(kind, normalize, x.detach(), src_ids, x.detach(), dst_ids, rel_emb.detach(),
                                         rel_ids, n_transform=n_transform, want_dst_vec=want_dst_vec)
        ctx.save_for_backward(x, src_ids, dst_ids, rel_emb, rel_ids)
        ctx.meta = (kind, normalize)
        return score, (tr if tr is not None else x.new_empty((0,))), (dv if dv is not None else x.new_empty((0,)))

    @staticmethod
    def backward(ctx, dscore, dtr, ddv):
        x, src_ids, dst_ids, rel_emb, rel_ids = ctx.saved_tensors
        kind, normalize = ctx.meta
        need_x, need_r = ctx.needs_input_grad[2], ctx.needs_input_grad[5]
        d_src, d_dst, d_rel = ops.score_bwd(kind, normalize, x, src_ids, x, dst_ids, rel_emb, rel_ids,
                                            dscore, dtr, want_src=need_x, want_dst=need_x, want_rel=need_r)
        dx = drel = None
        if need_x:
            if ddv is not None:
                d_dst = d_dst + ddv
            dx, keys = ops.index_add_sorted(torch.cat([d_src, d_dst], 0), torch.cat([src_ids, dst_ids], 0), x.size(0),
                                            return_keys=True)
            mark_sparse_rows(dx, keys)
        if need_r:
            drel = ops.index_add_sorted(d_rel, rel_ids, rel_emb.size(0))
        return None, None, dx, None, None, drel, None, None, None


class ScoreRowsFunction(torch.autograd.Function):
    """Module-level scorer call on already gathered rows (reference scorer.py:58-84, 154-186)
    and/or the relation operator ``transform`` (scorer.py:86-94, 188-201).  ``want_transform``: True = every
    triple, an int = the first that many triples."""

    @staticmethod
    def forward(ctx, kind, normalize, src_emb, dst_emb, rel_emb, rel_ids, want_score, want_transform):
        ctx.set_materialize_grads(False)
        B = int(rel_ids.numel())
        xd = dst_emb if dst_emb is not None else src_emb
        score, tr, _, _ = ops.score_fwd(kind, normalize, src_emb.detach(), None, xd.detach(), None, rel_emb.detach(),
                                        rel_ids,
                                        n_transform=B if want_transform is True else int(want_transform))
        ctx.save_for_backward(src_emb, xd, rel_emb, rel_ids)
        ctx.meta = (kind, normalize, dst_emb is not None)
        ctx.rel_sort = presort_on_side_stream(rel_ids, rel_emb.size(0)) if rel_emb.requires_grad and B else None
        empty = src_emb.new_empty((0,))
        return (score if want_score else empty), (tr if tr is not None else empty)

    @staticmethod
    def backward(ctx, dscore, dtr):
        src_emb, xd, rel_emb, rel_ids = ctx.saved_tensors
        kind, normalize, has_dst = ctx.meta
        d_src, d_dst, d_rel = ops.score_bwd(kind, normalize, src_emb, None, xd, None, rel_emb, rel_ids, dscore, dtr,
                                            want_src=ctx.needs_input_grad[2],
                                            want_dst=has_dst and ctx.needs_input_grad[3],
                                            want_rel=ctx.needs_input_grad[4])
        drel = None
        if d_rel is not None:
            pre = None
            if ctx.rel_sort is not None:
                keys, perm, ready = ctx.rel_sort
                torch.cuda.current_stream(d_rel.device).wait_event(ready)
                pre = (keys, perm)
            drel = ops.index_add_sorted(d_rel, rel_ids, rel_emb.size(0), presorted=pre)
        return None, None, d_src, d_dst, drel, None, None, None


class SplitLinearFunction(torch.autograd.Function):
    """y = x · Wᵀ for a bias-free ``nn.Linear`` on the tcgen05 GEMM (fp32 operands carried as bf16
    hi/lo planes in "fp32" mode).  Used for the ProjectionHead's linears over all N rows (reference
    core/model/projection.py:48-72, called from model.py:289-290), where torch's fp32 SGEMM would cost
    more than the whole GAT stack."""

    @staticmethod
    def forward(ctx, x, W, precision):
        with_lo = precision == "fp32"
        x2 = x.reshape(-1, x.size(-1))
        xp = ops.split_bf16(x2.detach(), with_lo)
        Wp = ops.split_bf16(W.detach(), with_lo)
        y = ops.gemm(xp, False, Wp, False, x2.size(0), W.size(0), W.size(1))
        ctx.save_for_backward(W)
        ctx.xp, ctx.with_lo, ctx.x_shape = xp, with_lo, x.shape
        return y.view(*x.shape[:-1], W.size(0))

    @staticmethod
    def backward(ctx, gy):
        (W,) = ctx.saved_tensors
        with_lo = ctx.with_lo
        g2 = gy.reshape(-1, gy.size(-1)).contiguous()
        M, N_out, K_in = g2.size(0), W.size(0), W.size(1)
        gp = ops.split_bf16(g2, with_lo)
        dW = dx = None
        if ctx.needs_input_grad[1]:  # dW[N_out, K_in] = gyᵀ · x  (both operands MN-major, split-K over rows)
            dW = ops.gemm(gp, True, ctx.xp, True, N_out, K_in, M, splits_k=ops.pick_splits_k(N_out, K_in, M, gy.device))
        if ctx.needs_input_grad[0]:  # dx[M, K_in] = gy · W  (B = Wᵀ stored K-major)
            WTp = ops.split_bf16(W.detach().t().contiguous(), with_lo)
            dx = ops.gemm(gp, False, WTp, False, M, K_in, N_out).view(ctx.x_shape)
        ctx.xp = None
        return dx, dW, None


def split_linear(x, weight, precision="fp32"):
    return SplitLinearFunction.apply(x, weight, precision)


class MarginLossFunction(torch.autograd.Function):
    """loss = mean_{b,k} relu(margin + neg[b,k] - pos[b]) on the flat score vector (one kernel that also
    produces d loss / d score; reference relgat_loss.py:51-54 + trainer score split)."""

    @staticmethod
    def forward(ctx, score, B, K, margin, projection_layout):
        loss, dscore = ops.margin_loss(score.detach(), B, K, margin, projection_layout)
        ctx.save_for_backward(dscore)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dscore,) = ctx.saved_tensors
        return dscore * g, None, None, None, None


def fused_margin_loss(score, num_pos: int, num_neg: int, margin: float, projection_path: bool = False):
    return MarginLossFunction.apply(score, int(num_pos), int(num_neg), float(margin), bool(projection_path))


class RankLossFunction(torch.autograd.Function):
    """Ranking loss of reference core/loss/relgat_loss.py:32-71 on pos [B], neg [B, K] (any strides) as one kernel
    that also yields d loss / d score; ``sanitize`` folds the trainer's nan_to_num (trainer:584, 647-648) in."""

    @staticmethod
    def forward(ctx, pos, neg, kind, margin, alpha, sanitize):
        loss, dpos, dneg = ops.rank_loss(pos.detach(), neg.detach(), kind, margin, alpha, sanitize)
        ctx.save_for_backward(dpos, dneg)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        dpos, dneg = ctx.saved_tensors
        return dpos * g, dneg * g, None, None, None, None


def fused_rank_loss(pos, neg, kind: str, margin: float = 1.0, alpha: float = 1.0, sanitize: bool = False):
    return RankLossFunction.apply(pos, neg, kind, float(margin if margin is not None else 0.0),
                                  float(alpha if alpha is not None else 1.0), bool(sanitize))


class ReconLossFunction(torch.autograd.Function):
    """Weighted reconstruction terms of reference core/loss/multi_objective_loss.py:62-80:
    w_pos*CosineLoss(f_r(A), B) + w_neg*(1 - CosineLoss(f_r(A), negB)) + w_mse*MSE(f_r(A), B), one kernel producing
    the three values and the gradients of the weighted sum.  Returns (weighted sum, values [3] detached)."""

    @staticmethod
    def forward(ctx, tr, dst, negdst, w_pos, w_neg, w_mse):
        has_neg = negdst is not None and negdst.numel() > 0
        values, d_tr, d_dst, d_neg = ops.recon_loss(tr.detach(), dst.detach(), negdst.detach() if has_neg else None,
                                                    w_pos, w_neg if has_neg else 0.0, w_mse)
        ctx.save_for_backward(d_tr, d_dst, d_neg)
        ctx.has_neg = has_neg
        ctx.mark_non_differentiable(values)
        total = values.new_zeros(())
        if w_pos != 0.0:  # zero-weight terms are dropped, as the reference does (multi_objective_loss.py:62-80)
            total = total + w_pos * values[0]
        if w_neg != 0.0:  # the mean over an empty negative set is nan in the reference too
            total = total + w_neg * (1.0 - values[1])
        if w_mse != 0.0:
            total = total + w_mse * values[2]
        return total, values

    @staticmethod
    def backward(ctx, g, _gv):
        d_tr, d_dst, d_neg = ctx.saved_tensors
        return d_tr * g, d_dst * g, (d_neg * g if ctx.has_neg else None), None, None, None


def fused_recon_loss(transformed_src, dst_vec, neg_dst_vec, w_pos: float, w_neg: float, w_mse: float):
    return ReconLossFunction.apply(transformed_src, dst_vec, neg_dst_vec, float(w_pos), float(w_neg), float(w_mse))


class GeluLayerNormFunction(torch.autograd.Function):
    """LayerNorm(GELU(h)) of a ProjectionHead hidden block (reference core/model/projection.py:56-62) as one kernel
    per direction instead of the two ATen element-wise passes."""

    @staticmethod
    def forward(ctx, h, gamma, beta, eps):
        shape = h.shape
        h2 = h.reshape(-1, shape[-1])
        y, mean, rstd = ops.gelu_layernorm_fwd(h2.detach(), None if gamma is None else gamma.detach(),
                                               None if beta is None else beta.detach(), eps)
        ctx.save_for_backward(h2, gamma, mean, rstd)
        ctx.has_affine = gamma is not None
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        h2, gamma, mean, rstd = ctx.saved_tensors
        want = ctx.has_affine and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        dh, dgamma, dbeta = ops.gelu_layernorm_bwd(dy.reshape(h2.shape).contiguous(), h2, gamma, mean, rstd, want_params=want)
        return dh.view(ctx.shape), (dgamma if ctx.has_affine else None), (dbeta if ctx.has_affine else None), None


def gelu_layernorm(h, gamma, beta, eps: float = 1e-5):
    return GeluLayerNormFunction.apply(h, gamma, beta, float(eps))
