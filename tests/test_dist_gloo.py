"""world_size-2 (and 3) gloo tests of the destination-range partition plumbing on CPU: bounds,
stable bucketing, padded all-gather layout, collectives and the batch-row exchange.  The per-rank
edge arithmetic is done by the oracle here (the CUDA kernels need a GPU); what is under test is
relgat_projector_b200/dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import relgat_oracle as O
from relgat_projector_b200 import dist as RD


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _graph(seed=0, n=61, e=700, r=5, h=3, f=4):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, e); dst = rng.integers(0, n - 3, e); rel = rng.integers(0, r, e)
    dst[:150] = 7  # a heavy destination makes the edge-balanced bounds uneven
    P = rng.standard_normal((n, h, f)); A = rng.standard_normal((h, r, f)); beta = rng.standard_normal(r) * 0.1
    return n, r, h, f, src, dst, rel, P, A, beta


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, r, h, f, src, dst, rel, P, A, beta = _graph()
        ei = torch.from_numpy(np.stack([src, dst])); et = torch.from_numpy(rel)
        part = RD.DstPartition(ei, et, n, r, rank, world, balance="edges", build_index=False, mode="allgather")
        # bounds agree with the oracle rule; bucketing is stable and complete
        ref_b = O.partition_bounds_np(dst, n, world, "edges")
        assert part.bounds == [int(v) for v in ref_b]
        ref_bucket = O.bucket_edges_np(src, dst, rel, ref_b)[rank]
        assert np.array_equal(part.edge_ids.numpy(), ref_bucket[3])
        assert np.array_equal(part.local_dst.numpy() + part.lo, ref_bucket[1])
        # padded layout round trip
        ids = torch.arange(n)
        padded = part.to_padded(ids)
        assert padded.unique().numel() == n and int(padded.max()) < part.n_padded
        # forward: local rows of P -> all-gather -> oracle edge arithmetic on the local bucket
        P_all_true = torch.from_numpy(P.reshape(n, -1))
        P_loc = part.pad_rows(part.local_rows(P_all_true))
        P_all = RD.all_gather_rows(P_loc, world)
        assert torch.equal(P_all[padded], P_all_true)
        g = O.graph_index_np(part.local_src.numpy(), part.local_dst.numpy(), part.local_rel.numpy(), part.n_padded, r)
        out, _, _, _ = O.layer_forward_closed(P_all.numpy().reshape(part.n_padded, h, f), A, beta, g)
        out_loc = torch.from_numpy(out[:part.n_local].reshape(part.n_local, -1))
        gathered = RD.all_gather_rows(part.pad_rows(out_loc), world)[padded]
        whole, _, _, _ = O.layer_forward_closed(P, A, beta, O.graph_index_np(src, dst, rel, n, r))
        assert np.allclose(gathered.numpy(), whole.reshape(n, -1), rtol=0, atol=1e-12)
        # halo mode: only referenced rows travel; extended layout [own | halo] reproduces the same output
        hp = RD.DstPartition(ei, et, n, r, rank, world, balance="edges", build_index=False, mode="halo")
        assert hp.n_src == hp.n_local + hp.n_halo and hp.n_halo <= n - hp.n_local
        P_own = part.local_rows(P_all_true).contiguous()
        recv = RD._all_to_all_rows(P_own[hp.send_idx], hp.send_counts, hp.recv_counts)
        assert torch.equal(recv, P_all_true[hp.halo_ids])
        P_ext = torch.cat([P_own, recv])
        gh = O.graph_index_np(hp.local_src.numpy(), hp.local_dst.numpy(), hp.local_rel.numpy(), max(hp.n_src, 1), r)
        out_h, _, _, _ = O.layer_forward_closed(P_ext.numpy().reshape(hp.n_src, h, f), A, beta, gh)
        assert np.allclose(out_h[:hp.n_local].reshape(hp.n_local, -1), whole.reshape(n, -1)[hp.lo:hp.hi], rtol=0, atol=1e-12)
        # reverse exchange: halo gradient rows return to their owners
        back = RD._all_to_all_rows(recv * 0 + float(rank + 1), hp.recv_counts, hp.send_counts)
        assert back.shape[0] == hp.n_send
        # reduce-scatter of per-rank partial rows == slice of the global sum
        part_rows = torch.full((part.n_padded, 3), float(rank + 1), dtype=torch.float64)
        mine = RD.reduce_scatter_rows(part_rows, world)
        assert mine.shape == (part.max_rows, 3) and torch.all(mine == sum(range(1, world + 1)))
        # batch-row exchange: every rank ends with the same complete rows
        x_global = torch.from_numpy(np.random.default_rng(5).standard_normal((n, 6)))
        ids_b = torch.tensor([0, n - 1, 7, 7, 30, part.bounds[1]])  # identical on every rank
        rows = RD.ExchangeBatchRows.apply(part.local_rows(x_global).contiguous(), ids_b, part)
        assert torch.equal(rows, x_global[ids_b])
        # parameter-gradient bucket
        p1, p2 = torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))
        p1.grad = torch.full((3, 2), float(rank)); p2.grad = torch.arange(5.0) * (rank + 1)
        RD.allreduce_grads([p1, p2])
        assert torch.all(p1.grad == sum(range(world)))
        assert torch.equal(p2.grad, torch.arange(5.0) * sum(range(1, world + 1)))
        ret[rank] = "ok"
    except Exception as exc:  # noqa: BLE001
        import traceback
        ret[rank] = "".join(traceback.format_exception(exc))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_plumbing_gloo(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rnk in range(world):
        assert ret.get(rnk) == "ok", ret.get(rnk)


def test_partition_bounds_single_process():
    n, r, h, f, src, dst, rel, P, A, beta = _graph(seed=3)
    d = torch.from_numpy(dst)
    for world in (1, 2, 4, 8):
        for balance in ("edges", "nodes"):
            b = RD.partition_bounds(d, n, world, balance)
            assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(world))
            assert b == [int(v) for v in O.partition_bounds_np(dst, n, world, balance)]


def _fd_worker(rank, world, port, ret):
    """Each rank owns two temp files with distinct contents and hands their descriptors to every peer."""
    import tempfile
    os.environ["MASTER_PORT"] = str(port)
    from relgat_projector_b200 import peer as RP
    files = []
    for i in range(2):
        f = tempfile.TemporaryFile()
        f.write(f"rank{rank}-fd{i}".encode())
        f.flush()
        files.append(f)
    got = RP._exchange_fds([f.fileno() for f in files], world, rank, tag="unittest")
    seen = {}
    for g in range(world):
        if g == rank:
            assert got[g] == []
            continue
        for i, fd in enumerate(got[g]):
            with os.fdopen(fd, "rb") as fh:
                fh.seek(0)
                seen[(g, i)] = fh.read().decode()
    ret[rank] = seen


@pytest.mark.parametrize("world", [2, 3])
def test_peer_table_descriptor_exchange(world):
    """Host plumbing of the peer tables (peer.py:_exchange_fds): every rank receives every peer's file
    descriptors, in order, over AF_UNIX / SCM_RIGHTS — no GPU involved."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fd_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for rank in range(world):
        want = {(g, i): f"rank{g}-fd{i}" for g in range(world) if g != rank for i in range(2)}
        assert ret[rank] == want


def test_peer_table_stride_rounding():
    """Table strides: smallest row count whose byte size is a granularity multiple for every row size."""
    from relgat_projector_b200 import peer as RP
    t = RP.PeerTables.__new__(RP.PeerTables)
    t.granularity = 2 << 20
    rows = t.stride_rows(300_000, [3200, 16, 32])
    assert rows >= 300_000 and all((rows * b) % t.granularity == 0 for b in (3200, 16, 32))
    assert rows - 300_000 < 131072  # 2 MiB / 16 B rows
    t.granularity = 1
    assert t.stride_rows(7, [12]) == 7 and t.stride_rows(0, [4]) == 1
