"""torchrun --nproc-per-node G tools/check_dist.py : partitioned (G ranks, NCCL) vs single-GPU results.

Every rank builds the same seeded graph and replicated parameters; the partitioned step's loss,
local node rows and (all-reduced) parameter gradients must match the unpartitioned step on rank 0's
GPU within 1e-4 relative (the reduction order over ranks differs, so not bit-exact)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relgat_projector_b200 as R  # noqa: E402
from relgat_projector_b200 import dist as RD, loss as L, peer as RP, synthetic as S  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", lr)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    n, t, r, d_in, h, f, b, k = 6000, 40000, 9, 96, 4, 40, 128, 3
    kg = S.tensor_kg(n, t, r, d_in, seed=7, device=str(dev), skew=0.6)
    torch.manual_seed(3)
    model = R.RelGATModel(kg.node_emb, kg.edge_index, kg.edge_type, num_rel=r, scorer_type="distmult", gat_out_dim=f,
                          gat_heads=h, dropout=0.0, gat_num_layers=2).to(dev)
    with torch.no_grad():
        for lyr in model.gat_layers:
            lyr.rel_bias.normal_(0, 0.1)
    gen = torch.Generator().manual_seed(1)
    src, rel, dst = (x.to(dev) for x in S.sample_batch(kg.train_triples.cpu(), n, b, k, gen))
    rank_loss = L.RelGATLoss("margin", None, 1.0, None, {})

    if "--peer" in sys.argv:  # peer tables (mapped NVLink rows) instead of the halo exchange
        tables = RP.PeerTables(world, rank, dev)
        part = RP.PeerPartition(kg.edge_index, kg.edge_type, n, r, rank, world, tables, h, f, 2,
                                halo_bf16="--halo-bf16" in sys.argv)
        part.E_local = part.E_fwd
        prg = RP.PeerRelGAT(model, part, kg.node_emb[part.lo:part.hi])
    else:
        part = RD.DstPartition(kg.edge_index, kg.edge_type, n, r, rank, world)
        prg = RD.PartitionedRelGAT(model, part)
    scores = prg.scores(src, rel, dst)
    pos, neg = L.split_scores(scores, b, k)
    loss = rank_loss.prepare_scores_and_compute_loss(pos, neg)
    loss.backward()
    prg.finish_backward()
    got = {name: p.grad.clone() for name, p in model.named_parameters()}
    x_local = prg.node_repr_local().detach().clone()
    model.zero_grad(set_to_none=True)

    scores_ref, _, _ = model(src, rel, dst, transform_to_input_if_possible=False)
    pr, ng = L.split_scores(scores_ref, b, k)
    loss_ref = rank_loss.prepare_scores_and_compute_loss(pr, ng)
    loss_ref.backward()
    x_ref = model.get_node_repr()

    def rel_err(a, bb):
        return float((a - bb).abs().max() / bb.abs().max().clamp_min(1e-30))

    errs = {"loss": abs(float(loss) - float(loss_ref)), "x_local": rel_err(x_local, x_ref[part.lo:part.hi])}
    for name, p in model.named_parameters():
        errs[name] = rel_err(got[name], p.grad)
    worst = max(errs.values())
    print(f"rank {rank}/{world}: rows [{part.lo},{part.hi}) edges {part.E_local} worst rel err {worst:.2e} "
          f"({max(errs, key=errs.get)})", flush=True)
    if "--halo-bf16" in sys.argv:  # stated tolerance of the bf16 mode: 2e-2 on values, 2e-1 max-norm on gradients
        good = all(v < (2e-2 if k in ("loss", "x_local") else 2e-1) for k, v in errs.items())
    else:
        good = worst < 1e-4  # fp32 contract
    ok = torch.tensor([1 if good else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if int(ok.item()) != 1:
        raise SystemExit(1)
    if rank == 0:
        print("check_dist OK")


if __name__ == "__main__":
    main()
