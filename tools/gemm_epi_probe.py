"""Where the time of the large-output GEMMs (P = X·Wᵀ, dX = dP·W) goes: full kernel, epilogue without its global
stores (RELGAT_GEMM_EPI=1), main loop only (RELGAT_GEMM_EPI=2), per pipeline depth and CTA-pair mode.

    python tools/gemm_epi_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relgat_projector_b200 import ops  # noqa: E402

SHAPES = [("P0", 300_000, 800, 1024), ("P1/dX1", 300_000, 800, 800)]


def t_ms(fn, n=6):
    for _ in range(2):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    for name, M, N, K in SHAPES:
        a = ops.split_bf16(torch.randn((M, K), generator=g, device=dev), True)
        b = ops.split_bf16(torch.randn((N, K), generator=g, device=dev), True)
        out = torch.empty((M, N), device=dev)
        for cg in ("2", "1"):
            for stages in ("0", "2", "3"):
                for epi in ("0", "1", "2"):
                    os.environ["RELGAT_GEMM_CG"] = cg
                    os.environ.pop("RELGAT_GEMM_STAGES", None)
                    if stages != "0":
                        os.environ["RELGAT_GEMM_STAGES"] = stages
                    os.environ.pop("RELGAT_GEMM_EPI", None)
                    if epi != "0":
                        os.environ["RELGAT_GEMM_EPI"] = epi
                    ms = t_ms(lambda: ops.gemm(a, False, b, False, M, N, K, out=out))
                    print(f"{name:7s} cg={cg} stages={stages or 'max'} epi={epi}  {ms:.3f} ms  "
                          f"{2.0 * M * N * K * 3 / ms / 1e9:.0f} TFLOP/s", flush=True)
        del a, b, out
    for k in ("RELGAT_GEMM_CG", "RELGAT_GEMM_STAGES", "RELGAT_GEMM_EPI"):
        os.environ.pop(k, None)


if __name__ == "__main__":
    main()
