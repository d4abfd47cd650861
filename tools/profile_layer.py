"""Runs the edge kernels of one RelGAT layer at a named config's shapes (for ncu captures).

    python tools/profile_layer.py --config c2 --iters 3
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relgat_projector_b200 import ops, synthetic as S  # noqa: E402
from relgat_projector_b200.graph import GraphIndex  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--skew", type=float, default=0.0)
    ap.add_argument("--bf16", action="store_true", help="bf16 feature storage (P, G)")
    args = ap.parse_args()
    cfg = S.CONFIGS[args.config]
    dev = torch.device("cuda:0")
    kg = S.tensor_kg(cfg["N"], cfg["T"], cfg["R"], 8, seed=42, device="cuda:0", skew=args.skew)
    g = GraphIndex(kg.edge_index, kg.edge_type, cfg["N"], cfg["R"])
    H, F, N = cfg["H"], cfg["F"], cfg["N"]
    gen = torch.Generator(device=dev).manual_seed(0)
    P = torch.randn((N, H * F), generator=gen, device=dev)
    if args.bf16:
        P = P.bfloat16()
    A = torch.randn((H, cfg["R"], F), generator=gen, device=dev) / F ** 0.5
    beta = torch.randn((cfg["R"],), generator=gen, device=dev) * 0.1
    dY = torch.randn((N, H * F), generator=gen, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    for it in range(args.iters):
        ev[0].record()
        out, act, _, z, minv, bias = ops.edge_fwd(P, A, beta, g, H, F, want_act=True, apply_elu=True, act_lo=not args.bf16)
        ev[1].record()
        G, t, hsum = ops.edge_bwd_prep(dY, out, bias, H, F, apply_elu=True, g_bf16=args.bf16)
        ev[2].record()
        _, planes, dz = ops.edge_bwd_src(P, G, A, z, minv, t, g, H, F, want_fp32=False, want_planes=True,
                                         planes_lo=not args.bf16)
        ev[3].record()
        dA, dbeta = ops.edge_bwd_rel(P, dz, hsum, g, H, F)
        ev[4].record()
        ops.edge_bwd_src(P, G, A, z, minv, t, g, H, F, want_fp32=False, want_planes=True, planes_lo=not args.bf16,
                         want_ds=True)  # the training path: dS columns; RELGAT_SRC_V2=0 selects the first-generation kernel
        ev[5].record()
        if not args.bf16:  # third generation: bulk-copy ring, rows [dPa | dS] (RELGAT_SRC3_DEPTH = ring depth)
            ops.edge_bwd_src(P, G, A, z, minv, t, g, H, F, want_fp32=False, want_planes=True, want_ds=True, a_term=False)
        ev[6].record()
        torch.cuda.synchronize()
        print("iter", it, "fwd %.3f prep %.3f src %.3f rel %.3f src(dS) %.3f src3 %.3f ms"
              % tuple(ev[i].elapsed_time(ev[i + 1]) for i in range(6)))


if __name__ == "__main__":
    main()
