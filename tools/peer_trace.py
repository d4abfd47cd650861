"""torchrun --nproc-per-node G tools/peer_trace.py : phase timeline of one training step on the peer-table path
(config 2 per GPU), from CUDA events on the three streams.  RELGAT_PEER_BLOCKS / RELGAT_PEER_DEEP_OVERLAP apply."""
import os
import sys

os.environ["RELGAT_PEER_TRACE"] = "1"
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relgat_projector_b200 as R  # noqa: E402
from relgat_projector_b200 import loss as L, peer as RP, synthetic as S  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    cfg = S.CONFIGS["c2"]
    n, t = cfg["N"] * world, cfg["T"] * world
    kg = S.tensor_kg(n, t, cfg["R"], cfg["D_in"], seed=42, device=str(dev))
    part = RP.PeerPartition(kg.edge_index, kg.edge_type, n, cfg["R"], rank, world, RP.PeerTables(world, rank, dev),
                            cfg["H"], cfg["F"], cfg["L"])
    x0 = kg.node_emb[part.lo:part.hi].clone()
    kg.node_emb = None
    torch.manual_seed(42)
    model = R.RelGATModel(x0, kg.edge_index[:, :1], kg.edge_type[:1], num_rel=cfg["R"], scorer_type=cfg["scorer"],
                          gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=0.0, gat_num_layers=cfg["L"]).to(dev)
    prg = RP.PeerRelGAT(model, part, model.node_emb_fixed)
    opt = torch.optim.Adam(model.parameters(), lr=2e-4)
    gen = torch.Generator().manual_seed(42)
    b, k = cfg["B"], cfg["K"]
    batches = [tuple(x.to(dev) for x in S.sample_batch(kg.train_triples[:200_000].cpu(), n, b, k, gen)) for _ in range(4)]

    def step(i):
        opt.zero_grad(set_to_none=True)
        loss = L.fused_margin_ranking_loss(prg.scores(*batches[i % 4]), b, k, 1.0)
        loss.backward()
        prg.finish_backward()
        opt.step()

    for i in range(4):
        step(i)
    dist.barrier()
    torch.cuda.synchronize()
    RP.trace_reset()
    step(5)
    rows = RP.trace_report()
    if rank == 0:
        print(f"| phase boundary (rank 0 of {world}, blocks={part.blocks}) | stream | ms since step start | delta on stream |")
        print("|---|---|---|---|")
        last = {}
        for label, st, ms in rows:
            d = ms - last.get(st, 0.0)
            last[st] = ms
            print(f"| {label} | {st} | {ms:.3f} | {d:.3f} |")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
