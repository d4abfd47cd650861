"""Host-side (Python) cost of one training step: cProfile over a few steps, top functions by cumulative time.

    python tools/host_profile.py --config c2 [--receptive-field] [--steps 30]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relgat_projector_b200 as R  # noqa: E402
from relgat_projector_b200 import loss as L, synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--receptive-field", action="store_true")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--top", type=int, default=45)
    a = ap.parse_args()
    cfg = S.CONFIGS[a.config]
    dev = torch.device("cuda:0")
    kg = S.tensor_kg(cfg["N"], cfg["T"], cfg["R"], cfg["D_in"], seed=42, device="cuda:0")
    torch.manual_seed(42)
    model = R.RelGATModel(kg.node_emb, kg.edge_index, kg.edge_type, num_rel=cfg["R"], scorer_type=cfg["scorer"],
                          gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=0.0, gat_num_layers=cfg["L"],
                          project_to_input_size=cfg["proj"], projection_layers=2).to(dev).train()
    model.receptive_field = a.receptive_field
    opt = torch.optim.Adam(model.parameters(), lr=2e-4)
    gen = torch.Generator().manual_seed(42)
    b, k = cfg["B"], cfg["K"]
    batches = [tuple(t.to(dev) for t in S.sample_batch(kg.train_triples.cpu(), cfg["N"], b, k, gen)) for _ in range(4)]
    rank_loss = L.RelGATLoss("margin", None, 1.0, None, {})
    multi = L.MultiObjectiveRelLoss(relgat_loss=rank_loss, run_config={}) if cfg["proj"] else None

    def step(i):
        src, rel, dst = batches[i % 4]
        opt.zero_grad(set_to_none=True)
        _, _, loss, *_ = L.calculate_loss(model, src, rel, dst, b, rank_loss, multi)
        loss.backward()
        opt.step()

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(a.steps):
        step(i)
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"{a.steps} steps: host enqueue {t_enq / a.steps * 1e3:.3f} ms/step, wall {t_all / a.steps * 1e3:.3f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for i in range(a.steps):
        step(i)
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime")
    st.print_stats(a.top)


if __name__ == "__main__":
    main()
