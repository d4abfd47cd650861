"""torchrun --nproc-per-node 2 tools/peer_probe.py : bandwidth of reads from a peer table over NVLink
(bulk copy, random row gather through torch, the forward edge kernel with local vs remote sources)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relgat_projector_b200 import graph as G, ops, peer as RP  # noqa: E402


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    H, F, R = 4, 200, 50
    C = H * F
    tables = RP.PeerTables(world, rank, dev)
    rows = tables.stride_rows(300_000, [4 * C])
    t = tables.allocate([("P", rows, (C,), torch.float32)], tag="probe")["P"]
    t.local.normal_()
    tok = torch.zeros(1, device=dev)
    dist.all_reduce(tok)
    torch.cuda.synchronize()
    peer = t.whole[rows:2 * rows]
    local = t.local
    buf = torch.empty_like(local)
    gb = local.numel() * 4 / 1e9
    res = {}
    res["copy local->local GB/s"] = gb / timeit(lambda: buf.copy_(local)) * 1e3
    res["copy peer->local GB/s"] = gb / timeit(lambda: buf.copy_(peer)) * 1e3
    idx = torch.randint(0, rows, (300_000,), device=dev)
    res["index_select local rows GB/s"] = gb / timeit(lambda: local.index_select(0, idx)) * 1e3
    res["index_select peer rows GB/s"] = gb / timeit(lambda: peer.index_select(0, idx)) * 1e3
    out = torch.empty(300_000, C, device=dev)
    res["pull_rows local GB/s"] = gb / timeit(lambda: ops.pull_rows(t.whole, idx, out)) * 1e3
    res["pull_rows peer GB/s"] = gb / timeit(lambda: ops.pull_rows(t.whole, idx + rows, out)) * 1e3
    small = tables.allocate([("s", tables.stride_rows(1_350_000, [16]), (4,), torch.float32)], tag="probe2")["s"]
    idz = torch.randint(0, 1_350_000, (1_350_000,), device=dev) + small.stride_rows
    zo = torch.empty(1_350_000, 4, device=dev)
    res["pull 16-byte rows from peer ms"] = timeit(lambda: ops.pull_rows(small.whole, idz, zo))
    if "--direct" in sys.argv:  # the edge kernel gathering straight from the peer (needs RELGAT_PF_DIST=0)
        n, E = 300_000, 1_350_000
        g0 = torch.Generator(device="cpu").manual_seed(1)
        src = torch.randint(0, n, (E,), generator=g0).to(dev)
        dst = torch.randint(0, n, (E,), generator=g0).to(dev)
        rel = torch.randint(0, R, (E,), generator=g0).to(dev)
        A = torch.randn(H, R, F, device=dev) * 0.1
        for name, off in (("local", 0), ("peer", rows)):
            gi = G.GraphIndex(torch.stack([src + off, dst]), rel, n, R, num_src_nodes=world * rows, src_chunks=False)
            ms = timeit(lambda: ops.edge_fwd(t.whole, A, None, gi, H, F, apply_elu=True), 3)
            res[f"edge_fwd {name} sources ms"] = ms
            res[f"edge_fwd {name} gathered GB/s"] = E * C * 4 / 1e9 / ms * 1e3
    if rank == 0:
        print("PF_DIST", os.environ.get("RELGAT_PF_DIST"), {k: round(v, 2) for k, v in res.items()}, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
