"""Destination-range partitioned RelGAT across the GPUs of one NVLink/NVSwitch box (one process
per GPU, torch.distributed over NCCL).  The reference is single-device; this is new (SURVEY.md
§2.3, §8(e)).

Layout: rank g owns the contiguous destination range [bounds[g], bounds[g+1]) — balanced by
in-edge count — i.e. those rows of every layer's input X, projection P and output.  Edges are
bucketed (stably) by owner(dst); each rank builds its CSR/CSC over its own destinations with
*global* source ids remapped into the padded all-gather layout (owner * max_rows + local row).

Per layer, forward : P_local = X_local · Wᵀ (tcgen05)  ->  exchange(P)  ->  fused edge kernel
           backward: edge kernels on local edges produce partial dP for every source they touch
                     ->  reverse exchange(dP)  ->  local dW partial, dX_local
Two exchange modes: "halo" (default) moves only the rows a rank's edges actually reference
(all-to-all with per-peer row lists computed once per graph; on a G-way partition of a random graph
with E/N = 4.5 that is 43% of the rows at G = 8), "allgather" moves every row (all-gather forward,
reduce-scatter backward).
End of step: all-reduce of the GAT parameter gradients (one flat bucket).
Batch rows x[src_ids] / x[dst_ids] are exchanged with one all-reduce of a [2B', D] buffer.
"""
from __future__ import annotations

import json
import os
import sys
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import ops
from .graph import GraphIndex


# ---------------------------------------------------------------------------------------------
# partition (pure index arithmetic; works on CPU and CUDA tensors)
# ---------------------------------------------------------------------------------------------
def partition_bounds(dst: torch.Tensor, num_nodes: int, world: int, balance: str = "edges") -> List[int]:
    """Contiguous destination ranges per rank.  "edges": equal in-edge counts (up to one node);
    "nodes": equal node counts.  Same rule as oracle.partition_bounds_np."""
    if balance == "nodes":
        return [(num_nodes * g) // world for g in range(world + 1)]
    counts = torch.bincount(dst, minlength=num_nodes)
    cum = torch.cat([counts.new_zeros(1), torch.cumsum(counts, 0)])
    total = int(cum[-1].item())
    targets = torch.tensor([(total * g + world - 1) // world for g in range(1, world)], device=cum.device,
                           dtype=cum.dtype)
    inner = torch.searchsorted(cum, targets, right=False).tolist() if world > 1 else []
    b = [0] + [int(v) for v in inner] + [num_nodes]
    for i in range(1, len(b)):  # monotone
        b[i] = max(b[i], b[i - 1])
    return b


class DstPartition:
    """This rank's share of the message-passing graph."""

    def __init__(self, edge_index: torch.Tensor, edge_type: torch.Tensor, num_nodes: int, num_rel: int,
                 rank: int, world: int, balance: str = "edges", build_index: bool = True, mode: str = "halo"):
        self.rank, self.world, self.N, self.R = rank, world, int(num_nodes), int(num_rel)
        self.mode = mode
        src, dst = edge_index[0], edge_index[1]
        self.bounds = partition_bounds(dst, self.N, world, balance)
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.hi - self.lo
        self.max_rows = max(self.bounds[g + 1] - self.bounds[g] for g in range(world))
        self.n_padded = world * self.max_rows
        sel = torch.nonzero((dst >= self.lo) & (dst < self.hi)).flatten()  # ascending: stable bucketing
        self.edge_ids = sel
        self.local_dst = dst[sel] - self.lo
        self.local_rel = edge_type[sel]
        self.E_local = int(sel.numel())
        if mode == "halo":
            self._setup_halo(src[sel])
        else:
            self.local_src = self.to_padded(src[sel])
            self.n_src = self.n_padded
        self.graph: Optional[GraphIndex] = None
        if build_index:
            self.graph = GraphIndex(torch.stack([self.local_src, self.local_dst]), self.local_rel,
                                    max(self.n_local, 1), self.R, num_src_nodes=max(self.n_src, 1))

    def _setup_halo(self, src_global: torch.Tensor) -> None:
        """Extended source layout of this rank: [own rows | halo rows grouped by owner, ascending id].
        Peers learn which of their rows to send through one exchange of the id lists."""
        dev = src_global.device
        uniq = torch.unique(src_global)  # sorted
        own = self.owner_of(uniq)
        remote = uniq[own != self.rank]
        r_owner = own[own != self.rank]
        self.recv_counts = torch.bincount(r_owner, minlength=self.world).tolist()  # rows I receive per peer
        self.n_halo = int(remote.numel())
        self.n_src = self.n_local + self.n_halo
        self.halo_ids = remote  # sorted by id == grouped by owner (ranges are contiguous and ascending)
        # ids -> extended row: own -> id - lo ; remote -> n_local + rank among remote ids
        pos = torch.searchsorted(remote, src_global)
        is_own = (src_global >= self.lo) & (src_global < self.hi)
        self.local_src = torch.where(is_own, src_global - self.lo, self.n_local + pos)
        # tell every owner which of its rows I need
        need_lists = [remote[r_owner == g] for g in range(self.world)]
        send_lists = exchange_id_lists(need_lists, self.world, self.rank, dev)
        self.send_counts = [int(t.numel()) for t in send_lists]
        self.send_idx = (torch.cat(send_lists) - self.lo) if send_lists else remote.new_zeros(0)
        self.n_send = int(self.send_idx.numel())
        self.send_sorted = torch.sort(self.send_idx, stable=True) if self.n_send else None  # static per graph

    def owner_of(self, ids: torch.Tensor) -> torch.Tensor:
        edges = torch.tensor(self.bounds[1:-1], device=ids.device, dtype=ids.dtype)
        return torch.bucketize(ids, edges, right=True) if self.world > 1 else torch.zeros_like(ids)

    def to_padded(self, ids: torch.Tensor) -> torch.Tensor:
        """Global node id -> row of the padded all-gather layout [world * max_rows]."""
        own = self.owner_of(ids)
        starts = torch.tensor(self.bounds[:-1], device=ids.device, dtype=ids.dtype)
        return own * self.max_rows + (ids - starts[own])

    def local_rows(self, x_global: torch.Tensor) -> torch.Tensor:
        return x_global[self.lo:self.hi]

    def pad_rows(self, x_local: torch.Tensor) -> torch.Tensor:
        if x_local.size(0) == self.max_rows:
            return x_local
        pad = x_local.new_zeros((self.max_rows - x_local.size(0),) + tuple(x_local.shape[1:]))
        return torch.cat([x_local, pad], 0)


# ---------------------------------------------------------------------------------------------
# collectives (backend agnostic: NCCL on the GPUs, gloo in the CPU tests)
# ---------------------------------------------------------------------------------------------
def all_gather_rows(x_local_padded: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """[max_rows, C] per rank -> [world * max_rows, C] (rank-major)."""
    out = x_local_padded.new_empty((world * x_local_padded.size(0),) + tuple(x_local_padded.shape[1:]))
    if world == 1:
        out.copy_(x_local_padded)
    else:
        dist.all_gather_into_tensor(out, x_local_padded.contiguous(), group=group)
    return out


def reduce_scatter_rows(x_all: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """[world * max_rows, C] partial sums per rank -> this rank's summed [max_rows, C]."""
    rows = x_all.size(0) // world
    out = x_all.new_empty((rows,) + tuple(x_all.shape[1:]))
    if world == 1:
        out.copy_(x_all)
    elif dist.get_backend(group) == "gloo":  # gloo has no reduce_scatter: all_reduce then slice
        tmp = x_all.clone()
        dist.all_reduce(tmp, group=group)
        out.copy_(tmp[dist.get_rank(group) * rows:(dist.get_rank(group) + 1) * rows])
    else:
        dist.reduce_scatter_tensor(out, x_all.contiguous(), group=group)
    return out


def allreduce_grads(params: Sequence[torch.nn.Parameter], group=None) -> None:
    """Sum the per-rank partial gradients of the GAT parameters in one flat bucket."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def _all_to_all_rows(send: torch.Tensor, send_counts, recv_counts, group=None) -> torch.Tensor:
    """Variable-size row exchange: rows of ``send`` are grouped by destination rank."""
    world = len(send_counts)
    out = send.new_empty((int(sum(recv_counts)),) + tuple(send.shape[1:]))
    if world == 1:
        return out
    if dist.get_backend(group) == "gloo":  # no all_to_all on gloo: all-gather padded buffers, slice mine
        rank = dist.get_rank(group)
        counts = torch.tensor(send_counts, dtype=torch.int64)
        all_counts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(all_counts, counts, group=group)
        mx = max(int(c.sum()) for c in all_counts)
        pad = send.new_zeros((max(mx, 1),) + tuple(send.shape[1:]))
        pad[: send.size(0)] = send
        bufs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        off = 0
        for g in range(world):
            start = int(all_counts[g][:rank].sum())
            n = int(all_counts[g][rank])
            out[off:off + n] = bufs[g][start:start + n]
            off += n
        return out
    dist.all_to_all_single(out, send.contiguous(), output_split_sizes=list(recv_counts),
                           input_split_sizes=list(send_counts), group=group)
    return out


def exchange_id_lists(need_lists, world: int, rank: int, device) -> list:
    """need_lists[g] = ids this rank needs from rank g  ->  list over peers h of the ids h needs
    from this rank (one size exchange + one payload exchange, once per graph)."""
    if world == 1:
        return [need_lists[0].new_zeros(0)]
    need_counts = [int(t.numel()) for t in need_lists]
    cnt = torch.tensor(need_counts, dtype=torch.int64, device=device)
    theirs = _all_to_all_rows(cnt.view(world, 1), [1] * world, [1] * world).view(-1).tolist()
    payload = torch.cat(need_lists) if need_lists else cnt.new_zeros(0)
    got = _all_to_all_rows(payload, need_counts, theirs)
    out, off = [], 0
    for h in range(world):
        out.append(got[off:off + theirs[h]])
        off += theirs[h]
    return out


# ---------------------------------------------------------------------------------------------
# partitioned GAT stack (CUDA)
# ---------------------------------------------------------------------------------------------
class PartitionedStackFunction(torch.autograd.Function):
    """out_local = RelGAT stack over this rank's destinations.  args: x0_local [n_local, D_in],
    then per layer (W, A, beta).  Gradients returned for the parameters are this rank's PARTIAL
    sums (call allreduce_grads afterwards)."""

    @staticmethod
    def forward(ctx, x0_local, part: DstPartition, heads: int, out_dim: int, precision: str, x0_planes, *params):
        L = len(params) // 3
        H, F = heads, out_dim
        C = H * F
        g = part.graph
        with_lo = precision == "fp32"
        n_loc, world = part.n_local, part.world
        planes = x0_planes if x0_planes is not None else ops.split_bf16(x0_local, with_lo)
        saved = []
        out = None
        for l in range(L):
            W, A, beta = params[3 * l], params[3 * l + 1], params[3 * l + 2]
            d_in = W.size(1)
            Wp = ops.split_bf16(W.detach(), with_lo)
            if part.mode == "halo":
                P_all = torch.empty((part.n_src, C), dtype=torch.float32, device=x0_local.device)
                ops.gemm(planes, False, Wp, False, n_loc, C, d_in, out=P_all[:n_loc])
                if part.n_send or part.n_halo:  # NCCL all-to-all of exactly the rows the peers' edges reference
                    P_all[n_loc:] = _all_to_all_rows(P_all[:n_loc].index_select(0, part.send_idx),
                                                     part.send_counts, part.recv_counts)
            else:
                P_pad = torch.zeros((part.max_rows, C), dtype=torch.float32, device=x0_local.device) \
                    if n_loc < part.max_rows else torch.empty((part.max_rows, C), dtype=torch.float32,
                                                              device=x0_local.device)
                ops.gemm(planes, False, Wp, False, n_loc, C, d_in, out=P_pad[:n_loc])
                P_all = all_gather_rows(P_pad, world)  # NCCL all-gather over NVLink
            last = l == L - 1
            out, act, _, z, minv, bias = ops.edge_fwd(P_all, A.detach(), None if beta is None else beta.detach(), g,
                                                      H, F, want_act=not last, apply_elu=True, act_lo=with_lo)
            saved.append(dict(xp=planes, Wp=Wp, P=P_all, out=out, minv=minv, z=z, bias=bias, A=A.detach(),
                              d_in=d_in, has_beta=beta is not None))
            planes = act
        ctx.saved, ctx.part, ctx.cfg = saved, part, (H, F, L, with_lo)
        ctx.x0_needs_grad = bool(x0_local.requires_grad)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        H, F, L, with_lo = ctx.cfg
        part = ctx.part
        g = part.graph
        C = H * F
        n_loc, world = part.n_local, part.world
        grads = [None] * (3 * L)
        dY, owned, dX = grad_out.contiguous(), False, None
        for l in reversed(range(L)):
            s = ctx.saved[l]
            G, t, hsum = ops.edge_bwd_prep(dY, s["out"], s["bias"], H, F, apply_elu=(l < L - 1), inplace=owned)  # fp32 storage
            dP_part, _, dz = ops.edge_bwd_src(s["P"], G, s["A"], s["z"], s["minv"], t, g, H, F,
                                              want_fp32=True, want_planes=False)
            dA, dbeta = ops.edge_bwd_rel(s["P"], dz, hsum, g, H, F, want_dbeta=s["has_beta"])
            if part.mode == "halo":
                dP_loc = dP_part[:n_loc]
                if part.n_send or part.n_halo:  # halo rows go back to their owners, folded in a fixed order
                    back = _all_to_all_rows(dP_part[n_loc:].contiguous(), part.recv_counts, part.send_counts)
                    dP_loc = ops.index_add_sorted(back, part.send_idx, n_loc, out=dP_loc.contiguous(),
                                                  presorted=part.send_sorted)
            else:
                dP_loc = reduce_scatter_rows(dP_part, world)[:n_loc]  # sum over ranks of my sources' rows
            del dP_part
            dPp = ops.split_bf16(dP_loc, with_lo)
            d_in = s["d_in"]
            dW = ops.gemm(dPp, True, s["xp"], True, C, d_in, n_loc,
                          splits_k=ops.pick_splits_k(C, d_in, n_loc, dY.device))
            grads[3 * l], grads[3 * l + 1], grads[3 * l + 2] = dW, dA, dbeta
            if l > 0 or ctx.x0_needs_grad:
                dX = ops.gemm(dPp, False, s["Wp"], True, n_loc, d_in, C)
                dY, owned = dX, True
        ctx.saved = None
        return (dX if ctx.x0_needs_grad else None, None, None, None, None, None, *grads)


class ExchangeBatchRows(torch.autograd.Function):
    """rows[i] = x_global[ids[i]] where x is row-sharded: every rank fills the rows it owns and
    one all-reduce completes the buffer.  Backward: the (replicated) row gradients are folded,
    in a deterministic order, into this rank's rows only."""

    @staticmethod
    def forward(ctx, x_local, ids, part: DstPartition):
        mine = (ids >= part.lo) & (ids < part.hi)
        rows = x_local.new_zeros((ids.numel(), x_local.size(1)))
        loc = torch.nonzero(mine).flatten()
        rows[loc] = x_local[ids[loc] - part.lo]
        if part.world > 1:
            dist.all_reduce(rows)
        ctx.save_for_backward(ids, loc)
        ctx.part, ctx.n_local = part, x_local.size(0)
        return rows

    @staticmethod
    def backward(ctx, grad_rows):
        ids, loc = ctx.saved_tensors
        part = ctx.part
        dx = ops.index_add_sorted(grad_rows.contiguous()[loc], ids[loc] - part.lo, ctx.n_local)
        return dx, None, None


class PartitionedRelGAT:
    """Per-rank driver around a replicated ``RelGATModel``-style parameter set: holds this rank's
    rows of the frozen embeddings, the partition and its graph index."""

    def __init__(self, model, part: DstPartition, x0_local: Optional[torch.Tensor] = None):
        self.model, self.part = model, part
        self.layers = model._layers()
        self.gat_params = [p for lyr in self.layers for p in lyr.parameters()]
        self.x0_local = (x0_local if x0_local is not None else part.local_rows(model.node_emb_fixed)).contiguous()
        self._planes = ops.split_bf16(self.x0_local, with_lo=(model.precision == "fp32"))

    def node_repr_local(self) -> torch.Tensor:
        flat = []
        for lyr in self.layers:
            flat += list(lyr.kernel_params())
        return PartitionedStackFunction.apply(self.x0_local, self.part, self.layers[0].heads, self.layers[0].out_dim,
                                              self.model.precision, self._planes, *flat)

    def scores(self, src_ids, rel_ids, dst_ids):
        from .peer import check_partitioned_model_supported
        check_partitioned_model_supported(self.model)  # no silent divergence from the single-GPU model (dropout)
        x_local = self.node_repr_local()
        ids = torch.cat([src_ids, dst_ids])
        rows = ExchangeBatchRows.apply(x_local, ids, self.part)
        if self.model.project_to_input_size:  # row-wise head: project the gathered batch rows only
            rows = self.model.projection(rows)
        b = src_ids.numel()
        return self.model.scorer(rows[:b], rel_ids, rows[b:])

    def finish_backward(self) -> None:
        extra = list(self.model.projection.parameters()) if self.model.project_to_input_size else []
        allreduce_grads(self.gat_params + extra)


# ---------------------------------------------------------------------------------------------
# bench.py entry for --gpus N > 1 (launched under torchrun)
# ---------------------------------------------------------------------------------------------
NVLINK_PEER_COPY_GBS = 770.0  # measured peer copy per direction on this pool (B200_PROFILING.md); nominal 900


def bench_main(args, cfg, rank, world, local_rank, metric, unit, load_peaks, ClockSampler, cpu_fn=None) -> int:
    import relgat_projector_b200 as R
    from . import loss as L, synthetic as S

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    # default: weak scaling — the graph grows with the number of GPUs (per-GPU work fixed at the 1-GPU
    # config); `--config c4` (an explicitly named config): strong scaling on that fixed graph
    strong = bool(getattr(args, "config", None))
    n_nodes, n_trip = (cfg["N"], cfg["T"]) if strong else (cfg["N"] * world, cfg["T"] * world)
    locality = float(getattr(args, "locality", 0.0) or 0.0)
    kg = S.tensor_kg(n_nodes, n_trip, cfg["R"], cfg["D_in"], seed=42, device=str(dev), locality=locality, blocks=world)
    E = int(kg.edge_index.size(1))
    exchange = getattr(args, "exchange", "peer")
    if args.precision != "fp32" and exchange == "peer":
        exchange = "halo"
    if exchange == "peer":
        from . import peer as RP
        usable, why = RP.peer_tables_available(world, rank, dev)
        if not usable:  # still a GPU path: the NCCL halo exchange below
            if rank == 0:
                print(f"[bench] peer tables unavailable ({why}); using the NCCL halo exchange", file=sys.stderr)
            exchange = "halo"
    if exchange == "peer":  # rows of other ranks are pulled from NVLink-mapped peer tables (peer.py); no exchange step
        halo_bf16 = getattr(args, "halo", "bf16") == "bf16"
        part = RP.PeerPartition(kg.edge_index, kg.edge_type, n_nodes, cfg["R"], rank, world,
                                RP.PeerTables(world, rank, dev), cfg["H"], cfg["F"], cfg["L"], halo_bf16=halo_bf16)
        part.E_local = part.E_fwd
    else:
        part = DstPartition(kg.edge_index, kg.edge_type, n_nodes, cfg["R"], rank, world, mode=exchange)
    x0_local = kg.node_emb[part.lo:part.hi].clone()  # this rank's rows of the frozen embeddings
    kg.node_emb = None
    torch.cuda.empty_cache()
    torch.manual_seed(42)  # identical replicated parameters on every rank
    model = R.RelGATModel(x0_local, kg.edge_index[:, :1], kg.edge_type[:1], num_rel=cfg["R"],
                          scorer_type=cfg["scorer"], gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=0.0,
                          gat_num_layers=cfg["L"], precision=args.precision).to(dev)
    if exchange == "peer":
        prg = RP.PeerRelGAT(model, part, model.node_emb_fixed)
    else:
        prg = PartitionedRelGAT(model, part, x0_local=model.node_emb_fixed)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=2e-4)
    rank_loss = L.RelGATLoss("margin", None, 1.0, None, {})
    b, k = cfg["B"], cfg["K"]
    gen = torch.Generator().manual_seed(42)  # same batches on every rank
    trip_cpu = kg.train_triples[:200_000].cpu()
    host = [tuple(t.pin_memory() for t in S.sample_batch(trip_cpu, n_nodes, b, k, gen)) for _ in range(8)]
    devb = [tuple(t.to(dev) for t in hb) for hb in host]

    def step(src, rel, dst):
        opt.zero_grad(set_to_none=True)
        scores = prg.scores(src, rel, dst)
        loss = L.fused_margin_ranking_loss(scores, b, k, 1.0)
        loss.backward()
        prg.finish_backward()
        opt.step()
        return loss

    def timed(fn, steps):
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / max(steps, 1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # slowest rank
        return float(ms.item())

    warm = max(args.warmup, 3)
    for i in range(warm):
        step(*devb[i % 8])
    ops.LAUNCHES = 0
    with ClockSampler(local_rank) as clocks:
        ms = timed(lambda i: step(*devb[i % 8]), args.steps)
    launches = ops.LAUNCHES

    def e2e(i):
        src, rel, dst = (t.to(dev, non_blocking=True) for t in host[i % 8])
        return float(step(src, rel, dst).item())
    e2e(0)
    ms_e2e = timed(e2e, args.steps)
    e_local = torch.tensor([part.E_local], device=dev)
    e_all = [torch.zeros_like(e_local) for _ in range(world)]
    dist.all_gather(e_all, e_local)
    # NVLink roofline of the dominant multi-GPU kernel: the forward halo pull of one layer, all ranks at once
    pull_gbs = None
    if exchange == "peer" and part.n_halo_f > 0:
        src_tab = "Pb0" if part.halo_bf16 else None
        for _ in range(2):
            part.pull("P0", part.pull_f, export=src_tab)
        dist.barrier()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            part.pull("P0", part.pull_f, export=src_tab)
        e.record()
        torch.cuda.synchronize()
        t_pull = torch.tensor([s.elapsed_time(e) / 5], device=dev)
        dist.all_reduce(t_pull, op=dist.ReduceOp.MAX)
        pull_bytes = part.n_halo_f * cfg["H"] * cfg["F"] * (2 if part.halo_bf16 else 4)
        pull_gbs = (pull_bytes, float(t_pull.item()))
    cpu = cpu_fn() if (rank == 0 and cpu_fn is not None) else None
    dist.barrier()
    if rank == 0:
        C = cfg["H"] * cfg["F"]
        if exchange == "peer":
            par = (f"dst-range partition x{world}; halo rows pulled by one gather kernel from peer tables mapped over "
                   f"NVLink (no pack / collective / unpack); fwd by destination owner, bwd by source owner (no "
                   f"cross-rank sum of dP); all-reduce of parameter grads")
            extra = {"halo_rows_fwd_rank0": part.n_halo_f, "halo_rows_bwd_rank0": part.n_halo_b,
                     "local_rows_rank0": part.n_local,
                     "nvlink_bytes_per_layer_pass_rank0": max(part.n_halo_f, part.n_halo_b) * C * (2 if part.halo_bf16 else 4)}
        else:
            par = (f"dst-range partition x{world}, NCCL {'all-to-all of halo rows' if exchange == 'halo' else 'all-gather'} "
                   f"of P fwd / of dP bwd, all-reduce of parameter grads")
            extra = {"halo_rows_rank0": getattr(part, "n_halo", None), "local_rows_rank0": part.n_local,
                     "exchange_bytes_per_layer_per_rank": getattr(part, "n_halo", 0) * C * 4,
                     "allgather_equivalent_bytes": (world - 1) * part.max_rows * C * 4}
        line = {
            "metric": metric, "value": E / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None,
            "dtype": "fp32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}{'' if strong else ' x' + str(world)}: synthetic KG {n_nodes} nodes / {n_trip} triplets ({E} "
                                   f"message-passing edges) / {cfg['R']} relations, {cfg['D_in']}-d, {cfg['L']} layers, "
                                   f"{cfg['H']} heads, gat-out-dim {cfg['F']}, {cfg['scorer']}, batch {b}, num-neg {k}",
                       "parallelism": par, "exchange": exchange,
                       **({"locality": locality, "note": "SUPPLEMENTARY graph with block locality, not the headline workload"}
                          if locality > 0 else {}),
                       "edges_per_rank": [int(t.item()) for t in e_all], **extra,
                       "l2": "inputs_exceed_L2"},
            "clocks": clocks.summary(),
            "e2e": {"value": E / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 3 * b * (1 + k) * 8, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": None, "cpu_baseline": None,
        }
        if pull_gbs is not None:
            pb, pt = pull_gbs
            gbs = pb / (pt * 1e-3) / 1e9
            line["roofline"] = {
                "kernel": "pull_rows (halo rows of one layer pass, rank with the slowest pull)", "bound": "nvlink",
                "achieved": round(gbs, 1), "peak": NVLINK_PEER_COPY_GBS, "unit": "GB/s",
                "frac": round(gbs / NVLINK_PEER_COPY_GBS, 4), "traffic": None,
                "bytes_per_launch": int(pb), "ms_per_launch": round(pt, 4),
                "peak_source": "measured peer copy per direction on this pool (B200_PROFILING.md: 770 GB/s; nominal 900)"}
            line["config"]["halo_rows"] = ("bf16 over NVLink, fp32 in local memory (stated tolerance 2e-2 relative; "
                                           "--halo fp32 for fp32 rows)" if part.halo_bf16 else "fp32")
        if cpu is not None:
            line["cpu_baseline"] = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
        ref_c4 = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                              "r02_bench_c4_n8_strong.json")
        if os.path.exists(ref_c4):  # north_star's 50 M-edge strong-scaling config, measured by the builder (not in this run)
            with open(ref_c4) as f:
                line["c4_strong_builder_run"] = json.load(f)
        print(json.dumps(line))
    dist.destroy_process_group()
    return 0
