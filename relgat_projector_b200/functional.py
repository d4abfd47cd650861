"""Autograd wiring of the RelGAT hot path over the sm_100a kernels.

``RelGATStackFunction`` runs L RelGAT layers (with the inter-layer ELU of reference
core/model/model.py:286-287 fused into the edge kernel's epilogue) as ONE autograd node:
forward = [split W -> tcgen05 GEMM -> fused edge kernel] per layer, backward = the closed
form of SURVEY.md §A.2 with the by-source / by-relation passes and the tcgen05 dW / dX GEMMs.
Nothing here does arithmetic in torch; torch allocates and orders the launches.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from .graph import GraphIndex

PRECISIONS = ("fp32", "bf16")

_SIDE_STREAMS = {}


def _side_stream(device) -> torch.cuda.Stream:
    """One extra stream per device: the HBM-bound by-relation pass runs beside the tensor-bound
    dW / dX GEMMs (they only share read-only inputs)."""
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
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


class RelGATStackFunction(torch.autograd.Function):
    """out = RelGAT_L(... ELU(RelGAT_1(x0)) ...) for layers sharing one graph.

    args: x0 [N, D_in] fp32, then per layer (W [H*F, D_in_l], A [H, R, F], beta [R] or None).
    ``x0_planes`` may carry a cached bf16 split of a frozen x0 (reference model.py:32 keeps the
    input embeddings as a buffer, so the split is paid once, not per step).
    """

    @staticmethod
    def forward(ctx, x0, graph: GraphIndex, heads: int, out_dim: int, precision: str, x0_planes, *params):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        if len(params) % 3 != 0 or not params:
            raise ValueError("params must be (W, A, beta) per layer")
        L = len(params) // 3
        H, F = heads, out_dim
        C = H * F
        N = graph.N
        with_lo = precision == "fp32"
        if x0.size(0) != N:
            raise ValueError(f"node_emb has {x0.size(0)} rows but the graph has {N} nodes")
        planes = x0_planes if x0_planes is not None else ops.split_bf16(x0, with_lo)
        saved = []
        out = None
        for l in range(L):
            W, A, beta = params[3 * l], params[3 * l + 1], params[3 * l + 2]
            d_in = W.size(1)
            if W.size(0) != C:
                raise ValueError(f"layer {l}: W must be [{C}, D_in]")
            Wp = ops.split_bf16(W.detach(), with_lo)
            # K-major copy of Wᵀ for dX = dP·W (3 MB transpose; the K-major B path is ~12% faster than MN-major)
            WTp = ops.split_bf16(W.detach().t().contiguous(), with_lo) if (l > 0 or x0.requires_grad) else None
            # "bf16": projected features are stored in bf16 (halves every gather of the edge kernels)
            P = ops.gemm(planes, False, Wp, False, N, C, d_in,
                         out_dtype=torch.float32 if with_lo else torch.bfloat16)
            last = l == L - 1
            out, act, _, z, minv, bias = ops.edge_fwd(P, A.detach(), None if beta is None else beta.detach(), graph,
                                                      H, F, want_act=not last, apply_elu=True, act_lo=with_lo)
            saved.append(dict(xp=planes, Wp=Wp, WTp=WTp, P=P, out=out, minv=minv, z=z, bias=bias, A=A.detach(),
                              d_in=d_in, has_beta=beta is not None))
            planes = act
        ctx.saved = saved
        ctx.graph = graph
        ctx.cfg = (H, F, L, with_lo)
        ctx.x0_needs_grad = bool(x0.requires_grad)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        H, F, L, with_lo = ctx.cfg
        g = ctx.graph
        C = H * F
        N = g.N
        grads: List[Optional[torch.Tensor]] = [None] * (3 * L)
        dY = grad_out.contiguous()
        nz_rows = sparse_rows_of(grad_out) if dY is grad_out else None
        owned = False
        dX = None
        for l in reversed(range(L)):
            s = ctx.saved[l]
            G, t, hsum = ops.edge_bwd_prep(dY, s["out"], s["bias"], H, F, apply_elu=(l < L - 1), inplace=owned,
                                           g_bf16=not with_lo, nonzero_rows=nz_rows if l == L - 1 else None)
            _, dPp, dz = ops.edge_bwd_src(s["P"], G, s["A"], s["z"], s["minv"], t, g, H, F,
                                          want_fp32=False, want_planes=True, planes_lo=with_lo)
            main = torch.cuda.current_stream(dY.device)
            side = _side_stream(dY.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):  # dA / dbeta: gather-bound, overlaps the GEMMs below
                dA, dbeta = ops.edge_bwd_rel(s["P"], dz, hsum, g, H, F, want_dbeta=s["has_beta"])
            d_in = s["d_in"]
            splits = ops.pick_splits_k(C, d_in, N, dY.device)
            dW = ops.gemm(dPp, True, s["xp"], True, C, d_in, N, splits_k=splits)
            grads[3 * l], grads[3 * l + 1], grads[3 * l + 2] = dW, dA, dbeta
            if l > 0 or ctx.x0_needs_grad:
                dX = ops.gemm(dPp, False, s["WTp"], False, N, d_in, C)
                dY, owned = dX, True
            main.wait_stream(side)  # join before any buffer of this layer is released or reused
            for tns in (dA, dbeta):
                if tns is not None:
                    tns.record_stream(main)
            del G, dPp, dz
        ctx.saved = None
        return (dX if ctx.x0_needs_grad else None, None, None, None, None, None, *grads)


def relgat_stack(x0, graph, heads, out_dim, layer_params: Sequence, precision="fp32", x0_planes=None):
    flat = []
    for W, A, beta in layer_params:
        flat += [W, A, beta]
    return RelGATStackFunction.apply(x0, graph, heads, out_dim, precision, x0_planes, *flat)


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
    and/or the relation operator ``transform`` (scorer.py:86-94, 188-201)."""

    @staticmethod
    def forward(ctx, kind, normalize, src_emb, dst_emb, rel_emb, rel_ids, want_score, want_transform):
        ctx.set_materialize_grads(False)
        B = int(rel_ids.numel())
        xd = dst_emb if dst_emb is not None else src_emb
        score, tr, _, _ = ops.score_fwd(kind, normalize, src_emb.detach(), None, xd.detach(), None, rel_emb.detach(),
                                        rel_ids, n_transform=B if want_transform else 0)
        ctx.save_for_backward(src_emb, xd, rel_emb, rel_ids)
        ctx.meta = (kind, normalize, dst_emb is not None)
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
        drel = ops.index_add_sorted(d_rel, rel_ids, rel_emb.size(0)) if d_rel is not None else None
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
