"""Host / device time of the per-step receptive-field block extraction (blocks.build_blocks) on a bench configuration.

usage: python tools/profile_blocks.py [--config c2] [--reps 20]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from relgat_projector_b200 import synthetic as S  # noqa: E402
from relgat_projector_b200.blocks import build_blocks  # noqa: E402
from relgat_projector_b200.graph import GraphIndex, StreamChunks  # noqa: E402


def wall(fn, reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    cfg = S.CONFIGS[a.config]
    dev = torch.device("cuda", 0)
    kg = S.tensor_kg(cfg["N"], cfg["T"], cfg["R"], cfg["D_in"], seed=42, device=str(dev))
    full = GraphIndex(kg.edge_index, kg.edge_type, cfg["N"], cfg["R"])
    gen = torch.Generator().manual_seed(1)
    src, rel, dst = (t.to(dev) for t in S.sample_batch(kg.train_triples.cpu(), cfg["N"], cfg["B"], cfg["K"], gen))
    ids = torch.cat([src, dst])
    ms, blk = wall(lambda: build_blocks(full, ids, cfg["L"]), a.reps)
    print(f"build_blocks: {ms:.3f} ms wall per call; blocks:",
          [(g.N_src, g.N, g.E, g.fwd_chunks.n_chunks, g.src_chunks.n_chunks) for g in blk.graphs])
    g = blk.graphs[0]
    ei = torch.stack([g.csr_src.long(), g.csr_dst.long()])
    et = g.csr_rel.long()
    ms, _ = wall(lambda: GraphIndex(ei, et, g.N, g.R, validate=False, num_src_nodes=g.N_src, degrees=False), a.reps)
    print(f"GraphIndex(first block): {ms:.3f} ms")
    ms, _ = wall(lambda: GraphIndex(ei, et, g.N, g.R, validate=False, num_src_nodes=g.N_src, degrees=False,
                                    fwd_chunks=False, src_chunks=False), a.reps)
    print(f"  without the chunk tables: {ms:.3f} ms")
    ms, _ = wall(lambda: StreamChunks(g.rowptr), a.reps)
    print(f"  StreamChunks(rowptr): {ms:.3f} ms")
    ms, _ = wall(lambda: StreamChunks(g.colptr), a.reps)
    print(f"  StreamChunks(colptr): {ms:.3f} ms")
    ms, _ = wall(lambda: g._build_rel_chunks(), a.reps)
    print(f"  rel chunks: {ms:.3f} ms")
    def both():
        gi = GraphIndex(ei, et, g.N, g.R, validate=False, num_src_nodes=g.N_src, degrees=False, fwd_chunks=False, src_chunks=False)
        t0 = time.perf_counter()
        a_ = StreamChunks(gi.rowptr)
        t1 = time.perf_counter()
        b_ = StreamChunks(gi.colptr)
        t2 = time.perf_counter()
        both.t = (t1 - t0, t2 - t1)
        return a_, b_
    ms, _ = wall(both, a.reps)
    print(f"  index + both chunk tables, called from outside: {ms:.3f} ms (last: {both.t[0]*1e3:.3f} + {both.t[1]*1e3:.3f})")
    ms, _ = wall(lambda: torch.unique(ids), a.reps)
    print(f"torch.unique(ids): {ms:.3f} ms")


if __name__ == "__main__":
    main()
