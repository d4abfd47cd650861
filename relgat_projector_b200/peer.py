"""Destination-range partition over peer tables: the multi-GPU path without an exchange step.

BASELINE.json's north_star partitions the graph by destination range and moves the transformed source
rows between GPUs over NVLink.  Here every rank keeps the rows other ranks need (projected features P,
output gradients G, per-node softmax statistics, raw logits) in *peer tables*
(``csrc/peer_table.cu``): one allocation per rank, all of them mapped back to back into every process.
A peer's rows are then ordinary addresses: loads travel over NVLink / NVSwitch, and there is no pack /
all-gather / unpack and no reduction of partial results.  One gather kernel per edge pass
(``relgat_pull_rows``) pulls the distinct rows a rank's edges reference into the tail of its own table
([own rows | pulled rows]) ahead of the unchanged edge kernel: gathering them in place from inside the
edge kernel would move every row once per EDGE instead of once per distinct row, and the kernels'
`prefetch.global.L2` on a peer address measured 70x slower than a plain load.

* forward:  a rank owns a destination range and the in-edges of those destinations; remote sources are
  pulled from the mapped P table.
* backward: a rank owns the SAME node range as sources and processes their out-edges (by-source pass
  and by-relation pass), pulling G / t / hsum / softmax statistics of the remote destinations and the
  saved logits from the mapped tables; for the last layer only the rows of the batch (dY is zero
  elsewhere) are written, pulled and cleared.  Every dP row is complete on its owner: no cross-rank sum.  Forward node
  rows are bit-identical to one GPU; gradients agree to rounding (a source's out-edges are summed in the
  order of the renumbered destinations).

Ordering across ranks is a stream-ordered NCCL all-reduce of one integer after each row block of a writer kernel
(``sync``): a reader kernel enqueued behind it cannot start before every rank's writer has finished.
Parameter gradients are per-rank partial sums, all-reduced once per step (``dist.allreduce_grads``).

``mode="sim"`` keeps all ranks' tables in one ordinary tensor of one process: the index logic and the
numerics of a W-rank run are then testable on a single GPU by stepping the W rank programs in
lock-step (``tests/test_gpu_model.py``)."""
from __future__ import annotations

import ctypes
import math
import os
import socket
import threading
import time
import types
from typing import Dict, Generator, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib, ops
from .dist import allreduce_grads, partition_bounds
from . import functional as _RF
from .functional import _side_stream, mark_sparse_rows, sparse_rows_of
from .graph import GraphIndex


# dW GEMMs on the side stream as well (beside the NVLink-bound halo pulls); 0 = only the by-relation pass
DEEP_OVERLAP = os.environ.get("RELGAT_PEER_DEEP_OVERLAP", "1") != "0"
# row blocks per writer kernel: the halo pull of block c runs (on its own stream) beside the GEMM / prep of block c+1.
# Measured on 2 GPUs (three sweeps): 4 blocks are 0.3-1.0 ms/step (2-5 %) faster than 1; the overlap is far from
# complete (the pull and the GEMM compete for the same SMs), see DESIGN.md section 6.
PIPELINE_BLOCKS = max(1, int(os.environ.get("RELGAT_PEER_BLOCKS", "4")))
# last layer's backward: dY is zero outside the batch rows, so only those rows of G / t / hsum are written, pulled
# and (afterwards) cleared — instead of moving a dense, almost-all-zero halo over NVLink
SPARSE_LAST = os.environ.get("RELGAT_PEER_SPARSE_LAST", "1") != "0"
# hidden layers' backward: the gradient rows of a hidden layer are exact zeros outside the in-neighbourhood of the
# batch's nodes (sources of the edges into them, and so on down).  Every rank holds the whole graph's in-edge lists, so
# it can tell which of ITS pulled rows are such zeros without asking their owners: those rows are not fetched (their
# local copies are kept at zero), the by-source pass still gathers every row.  Same arithmetic, ~1 % of the NVLink bytes
# of that pull.  Needs the sparse last layer (the batch ids).
NZ_PULL = os.environ.get("RELGAT_PEER_NZ_PULL", "1") != "0"
# halo rows cross NVLink rounded to bf16 (the owner exports a bf16 copy of its P / G rows; the pull widens them into
# the fp32 [own | pulled] table): half the link bytes of a layer pass.  Library default: off (fp32 rows, results equal
# to one GPU up to summation order); bench.py turns it on for the multi-GPU runs and states the tolerance (2e-2).
HALO_BF16_DEFAULT = os.environ.get("RELGAT_PEER_HALO", "fp32") == "bf16"
_COMM_STREAMS: Dict[int, "torch.cuda.Stream"] = {}

# poor man's timeline (nsys is not in the image): RELGAT_PEER_TRACE=1 records a CUDA event per phase boundary on the
# stream that runs the phase; trace_report() turns them into start / end offsets within the step
TRACE = os.environ.get("RELGAT_PEER_TRACE", "0") == "1"
_TRACE_EVENTS: List[Tuple[str, str, "torch.cuda.Event"]] = []


def _mark(label: str, stream_name: str = "main") -> None:
    if TRACE:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()  # on the current stream
        _TRACE_EVENTS.append((label, stream_name, ev))


def trace_reset() -> None:
    _TRACE_EVENTS.clear()


def trace_report() -> List[Tuple[str, str, float]]:
    """(label, stream, ms since the first mark) for every mark since trace_reset(); synchronises the device."""
    torch.cuda.synchronize()
    if not _TRACE_EVENTS:
        return []
    t0 = _TRACE_EVENTS[0][2]
    return [(label, st, t0.elapsed_time(ev)) for label, st, ev in _TRACE_EVENTS]


def _comm_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COMM_STREAMS[key]


# ---------------------------------------------------------------------------------------------
# peer tables
# ---------------------------------------------------------------------------------------------
class PeerTable:
    """``whole`` [world * stride_rows, *row_shape]: every rank's table, this rank's rows in
    ``local`` (a view).  ``slot_of[g]`` = position of rank g's table inside ``whole``."""

    def __init__(self, whole: torch.Tensor, stride_rows: int, slot_of: List[int], rank: int):
        self.whole, self.stride_rows, self.slot_of = whole, stride_rows, slot_of
        s = slot_of[rank]
        self.local = whole[s * stride_rows:(s + 1) * stride_rows]


class _DevicePointer:
    """Minimal __cuda_array_interface__ carrier: lets torch wrap a mapped range without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def _private_socket_dir() -> str:
    """A directory only this user can enter (mode 0700, ownership verified): the rendezvous sockets live there, so
    another local user can neither plant a socket under the expected name nor connect to ours."""
    base = os.path.join(os.environ.get("XDG_RUNTIME_DIR") or "/tmp", f"relgat_peer_{os.getuid()}")
    os.makedirs(base, mode=0o700, exist_ok=True)
    st = os.lstat(base)
    import stat
    if not stat.S_ISDIR(st.st_mode) or st.st_uid != os.getuid() or (st.st_mode & 0o077):
        raise RuntimeError(f"peer table exchange: {base} is not a private directory of this user")
    return base


def _peer_uid(conn: socket.socket) -> int:
    import struct
    cred = conn.getsockopt(socket.SOL_SOCKET, socket.SO_PEERCRED, struct.calcsize("3i"))
    _pid, uid, _gid = struct.unpack("3i", cred)
    return uid


def _exchange_fds(my_fds: List[int], world: int, rank: int, tag: str, timeout: float = 120.0) -> List[List[int]]:
    """Every rank hands its file descriptors to every peer over AF_UNIX sockets (SCM_RIGHTS).  Sockets live in a
    private per-user directory, a connecting peer must run under the same uid (SO_PEERCRED), and every blocking step
    has a timeout: a dead peer raises instead of hanging all ranks."""
    base = os.path.join(_private_socket_dir(), f"{os.environ.get('MASTER_PORT', '0')}_{tag}")
    path = f"{base}_{rank}.sock"
    if os.path.exists(path):
        os.unlink(path)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(world)
    srv.settimeout(timeout)
    errors: List[BaseException] = []

    def serve():
        try:
            served = 0
            while served < world - 1:
                conn, _addr = srv.accept()
                with conn:
                    conn.settimeout(timeout)
                    if _peer_uid(conn) != os.getuid():
                        continue  # not one of our ranks: no descriptors for it
                    socket.send_fds(conn, [b"f"], my_fds)
                    conn.recv(1)  # the peer confirms it holds the descriptors before we move on
                    served += 1
        except BaseException as exc:  # surfaced by the joining thread
            errors.append(exc)

    th = threading.Thread(target=serve, daemon=True)
    th.start()
    got: List[List[int]] = [[] for _ in range(world)]
    try:
        for g in range(world):
            if g == rank:
                continue
            peer_path = f"{base}_{g}.sock"
            deadline = time.time() + timeout
            while True:
                c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                try:
                    c.connect(peer_path)
                    break
                except (FileNotFoundError, ConnectionRefusedError):
                    c.close()
                    if time.time() > deadline:
                        raise RuntimeError(f"peer table exchange: rank {g} never opened {peer_path}")
                    time.sleep(0.05)
            with c:
                c.settimeout(timeout)
                _msg, fds, _flags, _addr = socket.recv_fds(c, 16, len(my_fds))
                if len(fds) != len(my_fds):
                    raise RuntimeError("peer table exchange: short descriptor list")
                got[g] = list(fds)
                c.send(b"k")
        th.join(timeout + 10.0)
        if th.is_alive() or errors:
            raise RuntimeError(f"peer table exchange: serving the peers failed ({errors[0] if errors else 'timeout'})")
    finally:
        srv.close()
        if os.path.exists(path):
            os.unlink(path)
    return got


class PeerTables:
    """Allocator of peer tables.  mode "vmm": real peer mapping (one process per GPU);
    mode "sim": all ranks' tables inside one tensor shared through ``sim_store`` (tests)."""

    def __init__(self, world: int, rank: int, device, mode: str = "vmm", sim_store: Optional[dict] = None):
        if mode not in ("vmm", "sim"):
            raise ValueError("mode must be 'vmm' or 'sim'")
        self.world, self.rank, self.device, self.mode = world, rank, torch.device(device), mode
        self.sim_store = sim_store if sim_store is not None else {}
        self._mapped: List[Tuple[int, int, int]] = []  # (base, bytes, handle)
        self._keep: List[object] = []
        if mode == "vmm":
            g = ctypes.c_ulonglong(0)
            _lib.check(_lib.load().relgat_peer_table_granularity(self._dev_index(), ctypes.byref(g)),
                       "relgat_peer_table_granularity")
            self.granularity = int(g.value)
            self.slot_of = [(o - rank) % world for o in range(world)]  # own table first
        else:
            self.granularity = 1
            self.slot_of = list(range(world))

    def _dev_index(self) -> int:
        return self.device.index if self.device.index is not None else torch.cuda.current_device()

    def stride_rows(self, min_rows: int, row_bytes: Sequence[int]) -> int:
        """Smallest row count >= min_rows whose byte size is a granularity multiple for every row size."""
        m = 1
        for b in row_bytes:
            m = math.lcm(m, self.granularity // math.gcd(self.granularity, b))
        return max(1, -(-max(min_rows, 1) // m)) * m

    def allocate(self, specs: Sequence[Tuple[str, int, Tuple[int, ...], torch.dtype]], tag: str = "t") -> Dict[str, PeerTable]:
        """specs: (name, stride_rows, row_shape, dtype).  Collective over all ranks in mode "vmm"."""
        out: Dict[str, PeerTable] = {}
        if self.mode == "sim":
            for name, stride, row_shape, dtype in specs:
                key = (tag, name)
                if key not in self.sim_store:
                    self.sim_store[key] = torch.zeros((self.world * stride, *row_shape), dtype=dtype, device=self.device)
                out[name] = PeerTable(self.sim_store[key], stride, self.slot_of, self.rank)
            return out
        lib = _lib.load()
        dev = self._dev_index()
        handles, fds, sizes = [], [], []
        for name, stride, row_shape, dtype in specs:
            nbytes = stride * int(torch.tensor([], dtype=dtype).element_size()) * int(math.prod(row_shape))
            if nbytes % self.granularity:
                raise ValueError(f"peer table {name}: {nbytes} bytes is not a multiple of {self.granularity}")
            h, fd = ctypes.c_ulonglong(0), ctypes.c_int(-1)
            rc = lib.relgat_peer_table_create(dev, nbytes, ctypes.byref(h), ctypes.byref(fd))
            self._check(rc, f"relgat_peer_table_create({name})")
            handles.append(int(h.value))
            fds.append(int(fd.value))
            sizes.append(nbytes)
        peer_fds = _exchange_fds(fds, self.world, self.rank, tag) if self.world > 1 else [[]]
        for i, (name, stride, row_shape, dtype) in enumerate(specs):
            arr = (ctypes.c_int * self.world)(*[(peer_fds[g][i] if g != self.rank else -1) for g in range(self.world)])
            base = ctypes.c_void_p(0)
            rc = lib.relgat_peer_table_map(dev, self.world, self.rank, handles[i], arr, sizes[i], ctypes.byref(base))
            self._check(rc, f"relgat_peer_table_map({name})")
            self._mapped.append((int(base.value), sizes[i], handles[i]))
            carrier = _DevicePointer(int(base.value), self.world * sizes[i])
            self._keep.append(carrier)
            raw = torch.as_tensor(carrier, device=self.device)
            whole = raw.view(dtype).view(self.world * stride, *row_shape)
            out[name] = PeerTable(whole, stride, self.slot_of, self.rank)
        for lst in peer_fds:
            for fd in lst:
                os.close(fd)
        for fd in fds:
            os.close(fd)
        return out

    @staticmethod
    def _check(rc: int, what: str) -> None:
        if rc != 0:
            drv = _lib.load().relgat_peer_table_last_driver_error()
            raise RuntimeError(f"{what} failed: rc={rc}, CUresult={drv} (peer tables need CUDA VMM with "
                               "POSIX file-descriptor handles and peer access between the GPUs)")

    def close(self) -> None:
        lib = _lib.load()
        for base, nbytes, handle in self._mapped:
            lib.relgat_peer_table_unmap(ctypes.c_void_p(base), self.world, nbytes, handle)
        self._mapped.clear()


def peer_tables_available(world: int, rank: int, device) -> Tuple[bool, str]:
    """Collective probe: maps one granule per rank end to end (create, export, descriptor hand-off, import, map,
    a remote read) and agrees on the outcome, so that every rank takes the same path afterwards."""
    ok, why = 1, ""
    try:
        tables = PeerTables(world, rank, device)
        rows = tables.stride_rows(1, [16])
        tb = tables.allocate([("probe", rows, (4,), torch.float32)], tag="probe")["probe"]
        tb.local.fill_(float(rank + 1))
        torch.cuda.synchronize()
    except Exception as exc:  # noqa: BLE001 - any failure means "use the NCCL exchange"
        ok, why, tables, tb = 0, f"{type(exc).__name__}: {exc}", None, None
    flag = torch.tensor([ok], device=device, dtype=torch.int32)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # also orders the fills before the remote reads below
    if int(flag.item()) == 1 and world > 1:
        nxt = (rank + 1) % world
        got = float(tb.whole[tables.slot_of[nxt] * tb.stride_rows, 0].item())
        flag.fill_(1 if got == float(nxt + 1) else 0)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1 and not why:
            why = "a read through the mapped range did not return the peer's data"
    if tables is not None:
        if world > 1:
            dist.barrier()
        del tb
        tables.close()
    return int(flag.item()) == 1, why


# ---------------------------------------------------------------------------------------------
# the partition: forward graph (in-edges of my destinations), backward graph (out-edges of my sources)
# ---------------------------------------------------------------------------------------------
class PeerIndexPlan:
    """Pure index arithmetic of one rank's share (torch ops on whatever device the edge list lives on — the CPU
    tests replay a whole partitioned layer from these arrays with the numpy oracle).

    A node table of rank g holds [own rows | pulled rows]; row v of the mapped range =
    ``slot_of[owner(v)] * stride_rows + (v - lo_owner)``.
      forward edges : in-edges of my destinations, sources renumbered into [own | forward halo]
      backward edges: out-edges of my sources, destinations renumbered into [own | backward halo]
    ``stride_fn(min_rows, row_bytes)`` rounds a table stride (PeerTables.stride_rows); ``agree_max`` makes all
    ranks agree on a maximum (an all-reduce in a real run)."""

    def __init__(self, edge_index: torch.Tensor, edge_type: torch.Tensor, num_nodes: int, rank: int, world: int,
                 slot_of: Sequence[int], stride_fn, row_bytes: Sequence[int], slot_bytes: Sequence[int],
                 blocks: int = 1, balance: str = "edges", agree_max=None):
        self.rank, self.world, self.N = rank, world, int(num_nodes)
        self.blocks = k_blocks = int(blocks)
        dev = edge_index.device
        src, dst = edge_index[0], edge_index[1]
        E = int(src.numel())
        self.bounds = partition_bounds(dst, self.N, world, balance)
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = n = self.hi - self.lo
        starts = torch.tensor(self.bounds[:-1], device=dev, dtype=torch.int64)
        inner = torch.tensor(self.bounds[1:-1], device=dev, dtype=torch.int64)
        sizes = torch.tensor([self.bounds[g + 1] - self.bounds[g] for g in range(world)], device=dev, dtype=torch.int64)
        slot_t = torch.tensor(list(slot_of), device=dev, dtype=torch.int64)

        def owner_of(ids):
            return torch.bucketize(ids, inner, right=True) if world > 1 else torch.zeros_like(ids)

        def halo_of(g: int):
            """(edges into g's range, their remote sources, edges out of g's range, their remote destinations)"""
            lo_g, hi_g = self.bounds[g], self.bounds[g + 1]
            in_f = (dst >= lo_g) & (dst < hi_g)
            in_b = (src >= lo_g) & (src < hi_g)
            hf = torch.unique(src[in_f & ~in_b])  # sorted ascending = grouped by owner
            hb = torch.unique(dst[in_b & ~in_f])
            return in_f, hf, in_b, hb

        in_f, halo_f, in_b, halo_b = halo_of(rank)
        self.halo_f, self.halo_b = halo_f, halo_b
        self.n_halo_f, self.n_halo_b = int(halo_f.numel()), int(halo_b.numel())
        need = n + max(self.n_halo_f, self.n_halo_b) + 1  # + one trash row behind the pulled rows (sparse pulls)
        if world > 1 and agree_max is not None:  # all ranks must agree on the table stride
            need = int(agree_max(need))
        elif world > 1:  # single process: look at every rank's halo
            for g in range(world):
                if g != rank:
                    _, hf, _, hb = halo_of(g)
                    need = max(need, self.bounds[g + 1] - self.bounds[g] + max(int(hf.numel()), int(hb.numel())) + 1)
        self.stride_rows = int(stride_fn(need, row_bytes))

        def row_id(ids):
            own = owner_of(ids)
            return slot_t[own] * self.stride_rows + (ids - starts[own])

        def pull_order(halo):
            """Position of every (sorted) halo id in the pulled rows, and the row count per pipeline block.
            Block c holds the rows that sit in block c of their OWNER's range (pulled as soon as every rank
            has written its block c).  Inside a block the owners are interleaved round-robin, starting at this
            rank's right-hand neighbour: sorted ids are grouped by owner, and all ranks sweeping their lists in
            that order would read from the same peer at the same time (measured: 240 GB/s per GPU on 8 GPUs
            instead of 660)."""
            own = owner_of(halo)
            blk = ((halo - starts[own]) * k_blocks) // sizes[own].clamp(min=1)  # row r of n is in block k*r // n
            grp = own * k_blocks + blk  # ascending along the sorted ids
            cnt = torch.bincount(grp, minlength=world * k_blocks)
            j = torch.arange(halo.numel(), device=dev) - (torch.cumsum(cnt, 0) - cnt)[grp]
            span = (int(j.max().item()) + 1 if halo.numel() else 1) * world
            perm = torch.argsort(blk * span + j * world + (own - rank - 1) % world)
            pos = torch.empty_like(perm)
            pos[perm] = torch.arange(halo.numel(), device=dev)
            per_block = torch.bincount(blk, minlength=k_blocks).tolist()
            return perm, pos, per_block

        def renumber(ids, halo, pos):
            mine = (ids >= self.lo) & (ids < self.hi)
            if halo.numel() == 0:
                return ids - self.lo
            at = torch.searchsorted(halo, ids).clamp_(max=halo.numel() - 1)
            return torch.where(mine, ids - self.lo, n + pos[at])

        self.owner_of, self.row_id = owner_of, row_id
        perm_f, pos_f, self.blk_f = pull_order(halo_f)
        perm_b, pos_b, self.blk_b = pull_order(halo_b)
        self.pos_b = pos_b  # position of every (sorted) backward-halo id among the pulled rows
        self.pull_f, self.pull_b = row_id(halo_f[perm_f]), row_id(halo_b[perm_b])  # rows of the mapped range to pull
        self.pull_gid_b = halo_b[perm_b]  # global node id of every pulled backward-halo row (pull order)
        # in-edges of EVERY node, by destination (int32): lets a rank work out, without asking anyone, which rows of a
        # hidden layer's gradient can be non-zero for a batch (sources of the edges into the batch's nodes, and so on
        # down) — the other rows are exact zeros at their owners and need not cross NVLink (see backward_steps)
        order = torch.argsort(dst, stable=True)
        self.in_src = src[order].to(torch.int32)
        self.in_rowptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev),
                                    torch.cumsum(torch.bincount(dst, minlength=self.N), 0)]).to(torch.int32)
        self.row_blocks = [(-(-c * n // k_blocks), -(-(c + 1) * n // k_blocks)) for c in range(k_blocks)]  # own rows per block
        # forward: in-edges of my destinations, original order (stable bucketing)
        sel_f = torch.nonzero(in_f).flatten()
        self.E_fwd = int(sel_f.numel())
        self.fwd_edges = (renumber(src[sel_f], halo_f, pos_f), dst[sel_f] - self.lo, edge_type[sel_f])
        # backward: out-edges of my sources
        sel_b = torch.nonzero(in_b).flatten()
        self.E_bwd = int(sel_b.numel())
        self.bwd_edges = (src[sel_b] - self.lo, renumber(dst[sel_b], halo_b, pos_b), edge_type[sel_b])
        # where the forward pass of the destination's owner stored the logit of each of my out-edges:
        # the owner's CSR order is the global stable by-destination order restricted to its range
        own_dst = owner_of(dst)
        per_owner = torch.bincount(own_dst, minlength=world)
        first = torch.cumsum(per_owner, 0) - per_owner
        order = torch.argsort(dst, stable=True)
        pos = torch.empty(E, dtype=torch.int64, device=dev)
        pos[order] = torch.arange(E, device=dev)
        self.stride_slots = int(stride_fn(int(per_owner.max().item()) if E else 1, slot_bytes))
        self.z_row_bwd = (slot_t[own_dst] * self.stride_slots + (pos - first[own_dst]))[sel_b]  # per backward edge


    def sparse_halo_targets(self, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """For batch node ids: (rows of the mapped range to read, rows of the own table to write) so that every batch
        node among this rank's backward halo lands in its pulled slot.  Ids that are not in the halo read own row 0
        into the trash row behind the pulled block (no host synchronisation, fixed list length)."""
        n, nh = self.n_local, self.n_halo_b
        if nh == 0:
            return torch.zeros_like(ids), torch.full_like(ids, n)
        at = torch.searchsorted(self.halo_b, ids).clamp_(max=nh - 1)
        hit = self.halo_b[at] == ids
        pos = self.pos_b[at]
        return torch.where(hit, self.pull_b[pos], torch.zeros_like(ids)), torch.where(hit, n + pos, torch.full_like(ids, n + nh))


class PeerPartition:
    """One rank's share on the GPU: the index plan, the two graph indexes built from it, and the peer tables."""

    def __init__(self, edge_index: torch.Tensor, edge_type: torch.Tensor, num_nodes: int, num_rel: int,
                 rank: int, world: int, tables: PeerTables, heads: int, out_dim: int, num_layers: int,
                 balance: str = "edges", tag: str = "p", blocks: Optional[int] = None,
                 halo_bf16: Optional[bool] = None):
        self.rank, self.world, self.N, self.R = rank, world, int(num_nodes), int(num_rel)
        self.H, self.F, self.L = heads, out_dim, num_layers
        self.halo_bf16 = HALO_BF16_DEFAULT if halo_bf16 is None else bool(halo_bf16)
        if self.halo_bf16 and (heads * out_dim) % 8 != 0:
            raise ValueError("bf16 halo rows need heads*out_dim to be a multiple of 8")
        self.tables = tables
        dev = edge_index.device
        C = heads * out_dim

        def agree_max(value: int) -> int:
            t = torch.tensor([value], device=dev, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return int(t.item())

        plan = PeerIndexPlan(edge_index, edge_type, num_nodes, rank, world, tables.slot_of, tables.stride_rows,
                             [4 * C, 2 * C, 4 * heads, 8 * heads], [4 * heads],
                             blocks=int(blocks if blocks is not None else PIPELINE_BLOCKS), balance=balance,
                             agree_max=agree_max if (world > 1 and tables.mode == "vmm") else None)
        self.plan = plan
        for name in ("bounds", "lo", "hi", "n_local", "blocks", "stride_rows", "stride_slots", "n_halo_f", "n_halo_b",
                     "pull_f", "pull_b", "blk_f", "blk_b", "row_blocks", "E_fwd", "E_bwd", "owner_of", "row_id",
                     "halo_b", "pos_b", "pull_gid_b"):
            setattr(self, name, getattr(plan, name))
        # (rowptr, sources) of the whole graph's in-edges in the shape ops.mark_sources expects
        self.in_graph = types.SimpleNamespace(rowptr=plan.in_rowptr.contiguous(), csr_src=plan.in_src.contiguous(),
                                              N=self.N, N_src=self.N)
        n = self.n_local
        fs, fd, fr = plan.fwd_edges
        self.fwd_graph = GraphIndex(torch.stack([fs, fd]), fr, max(n, 1), self.R,
                                    num_src_nodes=max(n + self.n_halo_f, 1), src_chunks=False)
        bs, bd, br = plan.bwd_edges
        self.bwd_graph = GraphIndex(torch.stack([bs, bd]), br, max(n + self.n_halo_b, 1), self.R,
                                    num_src_nodes=max(n, 1), fwd_chunks=False)
        self.z_index = plan.z_row_bwd[self.bwd_graph.csr_perm.long()].contiguous()  # by slot of the backward graph
        specs = []
        for l in range(num_layers):
            specs += [(f"P{l}", self.stride_rows, (C,), torch.float32), (f"G{l}", self.stride_rows, (C,), torch.float32),
                      (f"minv{l}", self.stride_rows, (heads, 2), torch.float32),
                      (f"t{l}", self.stride_rows, (heads,), torch.float32),
                      (f"hsum{l}", self.stride_rows, (heads,), torch.float32),
                      (f"z{l}", self.stride_slots, (heads,), torch.float32)]
            if self.halo_bf16:  # what the peers read: this rank's P / G rows rounded to bf16
                specs += [(f"Pb{l}", self.stride_rows, (C,), torch.bfloat16), (f"Gb{l}", self.stride_rows, (C,), torch.bfloat16)]
        specs.append(("out", self.stride_rows, (C,), torch.float32))
        self.t = tables.allocate(specs, tag=tag)
        self._token = torch.zeros(1, dtype=torch.int32, device=dev)
        self.batch_ids: Optional[torch.Tensor] = None  # node ids of the current batch (set by PeerBatchRows)
        self._sparse_clean = False  # last layer's G / t / hsum tables hold zeros outside the rows listed in _dirty
        self._dirty: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        # hidden layers, NZ_PULL: the pulled region of G{l} holds zeros outside the rows listed in _nz_dirty[l]
        self._nz_clean: Dict[int, bool] = {}
        self._nz_dirty: Dict[int, torch.Tensor] = {}

    def sparse_halo_targets(self, ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.plan.sparse_halo_targets(ids)

    def sync(self) -> None:
        """Stream-ordered rendezvous of all ranks (see module docstring)."""
        if self.tables.mode == "vmm" and self.world > 1:
            dist.all_reduce(self._token)

    def pull(self, name: str, ids: torch.Tensor, per_block: Optional[List[int]] = None, block: Optional[int] = None,
             export: Optional[str] = None):
        """Fills the pulled rows of table ``name`` (all of them, or those of one pipeline block) from their
        owners; returns the [own | pulled] view.  ``export``: read the owners' bf16 export table of that name instead
        (rows are widened to fp32 on arrival)."""
        tb, n = self.t[name], self.n_local
        k = int(ids.numel())
        a, b = 0, k
        if block is not None:
            a = sum(per_block[:block])
            b = a + per_block[block]
        ops.pull_rows(self.t[export].whole if export else tb.whole, ids[a:b], tb.local[n + a:n + b])
        return tb.local[:n + k]

    def export_rows(self, name: str, export: str, r0: int, r1: int) -> None:
        """bf16 copy of own rows [r0, r1) of table ``name`` into the export table the peers pull from."""
        if r1 > r0:
            ops.split_bf16(self.t[name].local[r0:r1], with_lo=False, out_hi=self.t[export].local[r0:r1])


# ---------------------------------------------------------------------------------------------
# the rank program: generators that yield where all ranks must have finished the previous phase
# ---------------------------------------------------------------------------------------------
def forward_steps(part: PeerPartition, planes, params: Sequence[torch.Tensor], with_lo: bool, saved: list,
                  x0_needs_grad: bool = False) -> Generator[None, None, torch.Tensor]:
    H, F, L = part.H, part.F, part.L
    C, n = H * F, part.n_local
    T = part.t
    out = None
    for l in range(L):
        W, A, beta = params[3 * l], params[3 * l + 1], params[3 * l + 2]
        d_in = W.size(1)
        Wp = ops.split_bf16(W.detach(), with_lo)
        WTp = ops.split_bf16(W.detach().t().contiguous(), with_lo) if (l > 0 or x0_needs_grad) else None
        main = torch.cuda.current_stream(W.device)
        comm = _comm_stream(W.device)
        P_ext = None
        _mark(f"fwd{l} start")
        for c, (r0, r1) in enumerate(part.row_blocks):
            if r1 > r0:
                ops.gemm(tuple(None if p is None else p[r0:r1] for p in planes), False, Wp, False, r1 - r0, C, d_in,
                         out=T[f"P{l}"].local[r0:r1])
                if part.halo_bf16:
                    part.export_rows(f"P{l}", f"Pb{l}", r0, r1)
            _mark(f"fwd{l} gemm block {c} done")
            yield  # every rank's block c of P is written
            _mark(f"fwd{l} rendezvous {c} done")
            comm.wait_stream(main)
            with torch.cuda.stream(comm):  # beside the GEMM of block c+1
                P_ext = part.pull(f"P{l}", part.pull_f, part.blk_f, c, export=f"Pb{l}" if part.halo_bf16 else None)
                _mark(f"fwd{l} pull block {c} done", "comm")
        main.wait_stream(comm)
        _mark(f"fwd{l} pulls joined")
        last = l == L - 1
        out, act, _, z, minv, bias = ops.edge_fwd(
            P_ext, A.detach(), None if beta is None else beta.detach(), part.fwd_graph, H, F,
            want_act=not last, apply_elu=True, act_lo=with_lo, out_buf=T["out"].local[:n] if last else None,
            z_out=T[f"z{l}"].local[:part.E_fwd], minv_out=T[f"minv{l}"].local[:n])
        _mark(f"fwd{l} edge kernel done")
        # third-generation by-source pass (rows [dPa | dS], dS·A folded into the GEMMs): parameter-only operands
        fold = None
        if _RF.SRC_V3 and _RF.USE_DS and with_lo and F % 4 == 0:
            fold = _RF.fold_operands(A.detach(), Wp, WTp, H, F, part.bwd_graph.R, d_in)
        saved.append(dict(xp=planes, Wp=Wp, WTp=WTp, out=out, bias=bias, A=A.detach(), d_in=d_in,
                          has_beta=beta is not None, fold=fold))
        planes = act
    return out


def backward_steps(part: PeerPartition, grad_out: torch.Tensor, saved: list, with_lo: bool,
                   x0_needs_grad: bool = False) -> Generator[None, None, Tuple[Optional[torch.Tensor], list]]:
    H, F, L = part.H, part.F, part.L
    C, n = H * F, part.n_local
    T, g = part.t, part.bwd_graph
    grads: List[Optional[torch.Tensor]] = [None] * (3 * L)
    dY, dX = grad_out.contiguous(), None
    keep: list = []
    # NZ_PULL: per hidden layer, the pull list with the entries of known-zero rows switched off (-1)
    nz_pull: Dict[int, torch.Tensor] = {}
    if (NZ_PULL and SPARSE_LAST and L > 1 and part.batch_ids is not None and part.n_halo_b > 0
            and sparse_rows_of(grad_out) is not None and dY is grad_out):
        bits = ops.mark_rows(part.batch_ids, part.N)
        gid = part.pull_gid_b
        for l in range(L - 2, -1, -1):
            bits = ops.mark_sources(bits, part.in_graph)  # rows of dL/d out_l that can be non-zero
            live = ((bits[gid >> 5] >> (gid & 31).to(torch.int32)) & 1).bool()
            nz_pull[l] = torch.where(live, part.pull_b, torch.full_like(part.pull_b, -1))
    for l in reversed(range(L)):
        s = saved[l]
        main = torch.cuda.current_stream(dY.device)
        comm = _comm_stream(dY.device)
        z = torch.empty((part.E_bwd, H), dtype=torch.float32, device=dY.device)
        comm.wait_stream(main)
        with torch.cuda.stream(comm):  # 4·H bytes per out-edge, from the logits' owners (written in the forward pass)
            ops.pull_rows(T[f"z{l}"].whole, part.z_index, z)
        _mark(f"bwd{l} start")
        rows_own = sparse_rows_of(grad_out) if (SPARSE_LAST and l == L - 1 and dY is grad_out) else None
        if rows_own is not None and part.batch_ids is not None:
            # dY is zero outside the batch rows: write / pull / clear only those rows of G, t, hsum
            nh = part.n_halo_b
            G_t, t_t, h_t = T[f"G{l}"], T[f"t{l}"], T[f"hsum{l}"]
            if not part._sparse_clean:
                G_t.local[:n + nh + 1].zero_()
                t_t.local[n:n + nh + 1].zero_()
                h_t.local[n:n + nh + 1].zero_()
                part._sparse_clean = True
            elif part._dirty is not None:  # rows the previous step wrote (every rank has long finished reading them)
                own_prev, halo_prev = part._dirty
                G_t.local.index_fill_(0, own_prev, 0.0)
                for tb in (G_t, t_t, h_t):
                    tb.local.index_fill_(0, halo_prev, 0.0)
            ops.edge_bwd_prep(dY, s["out"], s["bias"], H, F, apply_elu=False, G_out=G_t.local[:n],
                              t_out=t_t.local[:n], hsum_out=h_t.local[:n], nonzero_rows=rows_own)
            _mark(f"bwd{l} sparse prep done")
            yield  # every rank's batch rows of G / t / hsum are written
            _mark(f"bwd{l} rendezvous done")
            src_rows, dst_rows = part.sparse_halo_targets(part.batch_ids)
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                for tb in (G_t, t_t, h_t):
                    ops.pull_rows(tb.whole, src_rows, tb.local, out_ids=dst_rows)
                minv_ext = part.pull(f"minv{l}", part.pull_b)
                _mark(f"bwd{l} sparse pulls done", "comm")
            G_ext, t_ext, hsum_ext = G_t.local[:n + nh], t_t.local[:n + nh], h_t.local[:n + nh]
            part._dirty = (rows_own, dst_rows)
        else:
            if l == L - 1:
                part._sparse_clean = False
            ids_l = nz_pull.get(l)
            nh = part.n_halo_b
            if ids_l is not None:
                halo_G = T[f"G{l}"].local[n:n + nh]
                if not part._nz_clean.get(l):
                    halo_G.zero_()
                    part._nz_clean[l] = True
                elif l in part._nz_dirty:  # rows the previous step fetched (its by-source pass has long read them)
                    ops.zero_rows(halo_G, part._nz_dirty[l])
                pos = torch.arange(nh, device=dY.device)
                part._nz_dirty[l] = torch.where(ids_l >= 0, pos, torch.full_like(pos, -1))
            else:
                part._nz_clean[l] = False
            for c, (r0, r1) in enumerate(part.row_blocks):
                if r1 > r0:
                    ops.edge_bwd_prep(dY[r0:r1], s["out"][r0:r1], s["bias"][r0:r1], H, F, apply_elu=(l < L - 1),
                                      G_out=T[f"G{l}"].local[r0:r1], t_out=T[f"t{l}"].local[r0:r1],
                                      hsum_out=T[f"hsum{l}"].local[r0:r1],
                                      g_export=T[f"Gb{l}"].local[r0:r1] if part.halo_bf16 else None)
                _mark(f"bwd{l} prep block {c} done")
                yield  # every rank's block c of G / t / hsum is written
                _mark(f"bwd{l} rendezvous {c} done")
                comm.wait_stream(main)
                with torch.cuda.stream(comm):  # beside the prep of block c+1
                    G_ext = part.pull(f"G{l}", part.pull_b if ids_l is None else ids_l, part.blk_b, c,
                                      export=f"Gb{l}" if part.halo_bf16 else None)
                    t_ext, minv_ext, hsum_ext = (part.pull(f"{k}{l}", part.pull_b, part.blk_b, c) for k in ("t", "minv", "hsum"))
                    _mark(f"bwd{l} pull block {c} done", "comm")
        main.wait_stream(comm)
        _mark(f"bwd{l} pulls joined")
        P_loc = T[f"P{l}"].local[:n]
        use_ds = _RF.USE_DS
        fold = s.get("fold") if (use_ds and ops.src3_supported(P_loc, F) and ops.src3_supported(G_ext, F)) else None
        _, dPp, dz = ops.edge_bwd_src(P_loc, G_ext, s["A"], z, minv_ext, t_ext, g, H, F,
                                      want_fp32=False, want_planes=True, planes_lo=with_lo, want_ds=use_ds,
                                      a_term=fold is None)
        _mark(f"bwd{l} by-source kernel done")
        # dA / dbeta and dW are off the critical path (dX -> prep -> pull -> by-source pass of the layer below):
        # they run on the side stream, beside the NVLink-bound pulls, and are joined once at the end
        side = _side_stream(dY.device)
        side.wait_stream(main)
        d_in = s["d_in"]
        deep = DEEP_OVERLAP and l > 0  # the layer processed last has no pulls below it: dW on the main stream
        if use_ds:
            # widened rows [dP | dS]: one split-K GEMM gives dW and dS^T X; dA = (dS^T X) W^T; no
            # by-relation gather pass over P
            HR, Wd = H * g.R, dPp[0].size(1)

            def dw_and_tail():
                dW_ext = _RF.weight_grad_gemm(dPp, s["xp"], Wd, d_in, n, dY.device)
                Tp = ops.split_bf16(dW_ext[C:C + HR].contiguous(), with_lo)
                dW_ = dW_ext[:C]
                if fold is not None:  # rows are [dPa | dS]: dW = dPa^T X + A_bd^T (dS^T X)
                    dW_ = dW_ + ops.gemm(fold["Abd"], True, Tp, True, C, d_in, HR)
                dA_full = ops.gemm(Tp, False, s["Wp"], False, HR, C, d_in)
                dA_ = torch.stack([dA_full[h * g.R:(h + 1) * g.R, h * F:(h + 1) * F] for h in range(H)])
                return dW_, dA_, (ops.edge_bwd_beta(hsum_ext, g, H) if s["has_beta"] else None)

            if deep:
                with torch.cuda.stream(side):
                    dW, dA, dbeta = dw_and_tail()
                    _mark(f"bwd{l} dW done", "side")
            else:
                dW, dA, dbeta = dw_and_tail()
                _mark(f"bwd{l} dW done")
            dPc = tuple(None if p_ is None else p_[:, :C] for p_ in dPp)
        else:
            with torch.cuda.stream(side):
                dA, dbeta = ops.edge_bwd_rel(P_loc, dz, hsum_ext, g, H, F, want_dbeta=s["has_beta"])
                _mark(f"bwd{l} by-relation done", "side")
                if deep:
                    dW = ops.gemm(dPp, True, s["xp"], True, C, d_in, n, splits_k=ops.pick_splits_k(C, d_in, n, dY.device))
                    _mark(f"bwd{l} dW done", "side")
            if not deep:
                dW = ops.gemm(dPp, True, s["xp"], True, C, d_in, n, splits_k=ops.pick_splits_k(C, d_in, n, dY.device))
                _mark(f"bwd{l} dW done")
            dPc = dPp
        grads[3 * l], grads[3 * l + 1], grads[3 * l + 2] = dW, dA, dbeta
        if l > 0 or x0_needs_grad:
            if fold is not None:  # dX = [dPa | dS] · [W ; A_bd·W]
                dX = ops.gemm(dPp, False, fold["Bext"], False, n, d_in, dPp[0].size(1))
            else:
                dX = ops.gemm(dPc, False, s["WTp"], False, n, d_in, C)
            dY = dX
            _mark(f"bwd{l} dX done")
        keep.append((dPp, dz, z, hsum_ext))  # read on the side stream: released only after the join
        if not DEEP_OVERLAP:
            main.wait_stream(side)
    main = torch.cuda.current_stream(grad_out.device)
    main.wait_stream(_side_stream(grad_out.device))
    for tns in grads:
        if tns is not None:
            tns.record_stream(main)
    keep.clear()
    _mark("bwd side stream joined")
    return (dX if x0_needs_grad else None), grads


def drive(gen: Generator, sync) -> object:
    """Runs a rank program, calling ``sync`` at every yield."""
    try:
        while True:
            next(gen)
            sync()
    except StopIteration as stop:
        return stop.value


def drive_lockstep(gens: Sequence[Generator]) -> list:
    """Single-process simulation: advances all rank programs phase by phase."""
    results = [None] * len(gens)
    live = list(range(len(gens)))
    while live:
        for i in list(live):
            try:
                next(gens[i])
            except StopIteration as stop:
                results[i] = stop.value
                live.remove(i)
        if live and len(live) != len(gens):
            raise RuntimeError("rank programs disagree on the number of phases")
    return results


class PeerStackFunction(torch.autograd.Function):
    """out_local = RelGAT stack over this rank's destinations, sources read from the peer tables.
    Parameter gradients are this rank's partial sums (all-reduce them afterwards)."""

    @staticmethod
    def forward(ctx, x0_local, part: PeerPartition, precision: str, x0_planes, *params):
        with_lo = precision == "fp32"
        if not with_lo:
            raise ValueError("the peer-table path stores fp32 rows (precision='fp32')")
        planes = x0_planes if x0_planes is not None else ops.split_bf16(x0_local, with_lo)
        saved: list = []
        out = drive(forward_steps(part, planes, params, with_lo, saved, bool(x0_local.requires_grad)), part.sync)
        ctx.saved, ctx.part, ctx.with_lo = saved, part, with_lo
        ctx.x0_needs_grad = bool(x0_local.requires_grad)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        dX, grads = drive(backward_steps(ctx.part, grad_out, ctx.saved, ctx.with_lo, ctx.x0_needs_grad), ctx.part.sync)
        ctx.saved = None
        return (dX, None, None, None, *grads)


class PeerBatchRows(torch.autograd.Function):
    """rows[i] = x[ids[i]] for batch node ids anywhere in the graph: the last layer's output rows live in the
    ``out`` peer table, so the gather is a read of mapped rows (no collective).  Backward folds the row
    gradients (every rank computes all of them: the batch is replicated) into this rank's rows, in order.
    No host synchronisation: rows of other ranks are folded into scratch rows that are dropped."""

    @staticmethod
    def forward(ctx, x_local, ids, part: PeerPartition):
        n = part.n_local
        if x_local.data_ptr() != part.t["out"].local.data_ptr():  # the stack's last layer writes there directly
            part.t["out"].local[:n].copy_(x_local)
        part.sync()
        rows = x_local.new_empty((ids.numel(), x_local.size(1)))
        ops.pull_rows(part.t["out"].whole, part.row_id(ids), rows)
        mine = (ids >= part.lo) & (ids < part.hi)
        # rows of other ranks get one scratch row each (a shared scratch row would be one long serial segment)
        ctx.save_for_backward(torch.where(mine, ids - part.lo, n + torch.arange(ids.numel(), device=ids.device)),
                              torch.where(mine, ids - part.lo, torch.zeros_like(ids)))
        ctx.n_local = n
        part.batch_ids = ids
        return rows

    @staticmethod
    def backward(ctx, grad_rows):
        keys, rows_own = ctx.saved_tensors
        dx = ops.index_add_sorted(grad_rows.contiguous(), keys, ctx.n_local + keys.numel())[:ctx.n_local]
        mark_sparse_rows(dx, rows_own)  # own batch rows (row 0 stands in for rows of other ranks: harmless, see peer.py)
        return dx, None, None


class PeerRelGAT:
    """Per-rank driver around a replicated ``RelGATModel`` parameter set (same role as
    ``dist.PartitionedRelGAT``, peer tables instead of an exchange)."""

    def __init__(self, model, part: PeerPartition, x0_local: torch.Tensor):
        self.model, self.part = model, part
        self.layers = model._layers()
        self.gat_params = [p for lyr in self.layers for p in lyr.parameters()]
        self.x0_local = x0_local.contiguous()
        self._planes = ops.split_bf16(self.x0_local, with_lo=True)

    def node_repr_local(self) -> torch.Tensor:
        flat = []
        for lyr in self.layers:
            flat += list(lyr.kernel_params())
        return PeerStackFunction.apply(self.x0_local, self.part, self.model.precision, self._planes, *flat)

    def batch_rows(self, ids: torch.Tensor) -> torch.Tensor:
        """``model.batch_rows(ids)`` on the partitioned graph: the stack's rows of the batch nodes, through the
        projection head when the model has one (a row-wise map: only the gathered rows are projected)."""
        check_partitioned_model_supported(self.model)
        x_local = self.node_repr_local()
        rows = PeerBatchRows.apply(x_local, ids, self.part)
        return self.model.projection(rows) if self.model.project_to_input_size else rows

    def scores(self, src_ids, rel_ids, dst_ids):
        rows = self.batch_rows(torch.cat([src_ids, dst_ids]))
        b = src_ids.numel()
        return self.model.scorer(rows[:b], rel_ids, rows[b:])

    def finish_backward(self) -> None:
        extra = list(self.model.projection.parameters()) if self.model.project_to_input_size else []
        allreduce_grads(self.gat_params + extra)


def check_partitioned_model_supported(model) -> None:
    """The partitioned rank programs run the layers without dropout masks: refuse a training-mode model whose layer
    or projection dropout is active instead of silently training a different model than one GPU would (reference
    layer.py:296-297, 321-322; projection.py:69-72)."""
    if not model.training:
        return
    for lyr in model._layers():
        if lyr.dropout.p > 0.0 or lyr.rel_attn_drop.p > 0.0:
            raise NotImplementedError("partitioned (multi-GPU) training runs without dropout masks: construct the model "
                                      "with dropout=0 and relation_attn_dropout=0, or train on one GPU")
    if model.project_to_input_size and isinstance(model.projection.dropout, torch.nn.Dropout) and model.projection.dropout.p > 0:
        raise NotImplementedError("partitioned (multi-GPU) training does not support projection_dropout > 0")
