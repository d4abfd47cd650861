"""Times the tcgen05 GEMM on the four shapes of a config-2 step for a list of N-tile widths (RELGAT_GEMM_BN).

    python tools/gemm_sweep.py            # prints ms and useful TFLOP/s (3 passes counted) per shape and BN
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relgat_projector_b200 import ops  # noqa: E402

SHAPES = [  # (name, M, N, K, a_mn, b_mn, splits)
    ("P0 = X0 W0^T", 300_000, 800, 1024, False, False, 1),
    ("P1 = X1 W1^T / dX1", 300_000, 800, 800, False, False, 1),
    ("dW0 = dP^T X0", 800, 1024, 300_000, True, True, None),
    ("dW1 = dP^T X1", 800, 800, 300_000, True, True, None),
    ("dW0x = [dP|dS]^T X0", 1000, 1024, 300_000, True, True, None),
    ("dW1x = [dP|dS]^T X1", 1000, 800, 300_000, True, True, None),
    ("dW1x^T = X1^T [dP|dS]", 800, 1000, 300_000, True, True, None),
]


def planes(rows, cols, dev):
    g = torch.Generator(device=dev).manual_seed(rows + cols)
    x = torch.randn((rows, cols), generator=g, device=dev)
    return ops.split_bf16(x, True)


def main():
    dev = torch.device("cuda:0")
    bns = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "128", "160", "208", "256"])]
    for name, M, N, K, a_mn, b_mn, splits in SHAPES:
        a = planes(K if a_mn else M, M if a_mn else K, dev)
        b = planes(K if b_mn else N, N if b_mn else K, dev)
        for bn in bns:
            if bn:
                os.environ["RELGAT_GEMM_BN"] = str(bn)
            else:
                os.environ.pop("RELGAT_GEMM_BN", None)
            sk_list = [1] if splits == 1 else [ops.pick_splits_k(M, N, K, dev), 4, 8, 12, 16]
            for sk in dict.fromkeys(sk_list):
                for _ in range(2):
                    ops.gemm(a, a_mn, b, b_mn, M, N, K, splits_k=sk)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(5):
                    ops.gemm(a, a_mn, b, b_mn, M, N, K, splits_k=sk)
                e.record()
                torch.cuda.synchronize()
                ms = s.elapsed_time(e) / 5
                print(f"{name:22s} BN={bn or 'auto':>4} splits={sk:2d}  {ms:.3f} ms  {2.0 * M * N * K * 3 / ms / 1e9:.0f} TFLOP/s (3 passes)", flush=True)
        del a, b
        torch.cuda.empty_cache()
    # the dX GEMM with the backward prep of the layer below fused into its epilogue, against the unfused pair
    M, H, F, K = 300_000, 4, 200, 800
    N = H * F
    g = torch.Generator(device=dev).manual_seed(1)
    dP, WT = planes(M, K, dev), planes(N, K, dev)
    y = torch.randn((M, N), generator=g, device=dev)
    bias = torch.randn((M,), generator=g, device=dev)

    def t_ms(fn, n=5):
        for _ in range(2):
            fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    fused = t_ms(lambda: ops.gemm_dx_prep(dP, WT, M, N, K, y, bias, H, F, apply_elu=True))
    gemm_only = t_ms(lambda: ops.gemm(dP, False, WT, False, M, N, K))
    dX = ops.gemm(dP, False, WT, False, M, N, K)
    prep_only = t_ms(lambda: ops.edge_bwd_prep(dX, y, bias, H, F, apply_elu=True))
    print(f"dX + prep of the layer below: fused epilogue {fused:.3f} ms; unfused GEMM {gemm_only:.3f} + prep {prep_only:.3f} "
          f"= {gemm_only + prep_only:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
