"""Kernel timeline of one training step of the hot path (CUPTI via torch.profiler; nsys is not in the image).

    python tools/trace_step.py --config c2 --out gpurun_out/trace_step.md

Prints every kernel of ONE step (after warm-up) with its stream, start offset, duration and the idle gap since the
previous kernel on any stream, plus the busy/idle split of the step: the evidence for which kernels sit on the
critical chain, which overlap, and how much of the step is launch gaps.  Times under the profiler are a few percent
slower than bench.py's; use them for structure, not for headline numbers.
"""
import argparse
import json
import os
import re
import sys
import tempfile
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import relgat_projector_b200 as R  # noqa: E402
from relgat_projector_b200 import loss as L, synthetic as S  # noqa: E402


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*", "", name)
    name = name.replace("relgat::", "").replace("at::native::", "").replace("<unnamed>::", "")
    return name[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--receptive-field", action="store_true", help="run the step on the batch's blocks (blocks.py)")
    ap.add_argument("--out", default="gpurun_out/trace_step.md")
    args = ap.parse_args()
    cfg = S.CONFIGS[args.config]
    dev = torch.device("cuda:0")
    kg = S.tensor_kg(cfg["N"], cfg["T"], cfg["R"], cfg["D_in"], seed=42, device="cuda:0")
    torch.manual_seed(42)
    model = R.RelGATModel(kg.node_emb, kg.edge_index, kg.edge_type, num_rel=cfg["R"], scorer_type=cfg["scorer"],
                          gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=args.dropout, gat_num_layers=cfg["L"],
                          project_to_input_size=cfg["proj"], projection_layers=2, precision=args.precision).to(dev)
    model.train()
    model.receptive_field = args.receptive_field
    opt = torch.optim.Adam(model.parameters(), lr=2e-4)
    b, k = cfg["B"], cfg["K"]
    gen = torch.Generator().manual_seed(42)
    batches = [tuple(t.to(dev) for t in S.sample_batch(kg.train_triples.cpu(), cfg["N"], b, k, gen)) for _ in range(4)]
    rank_loss = L.RelGATLoss("margin", None, 1.0, None, {})
    multi_loss = L.MultiObjectiveRelLoss(relgat_loss=rank_loss, run_config={}) if cfg["proj"] else None

    def step(i):
        src, rel, dst = batches[i % 4]
        opt.zero_grad(set_to_none=True)
        _, _, loss, *_ = L.calculate_loss(model, src, rel, dst, b, rank_loss, multi_loss)
        loss.backward()
        opt.step()

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(3):
            step(i)
            torch.cuda.synchronize()  # a clear gap between steps (the idle time right after it is launch latency)
            time.sleep(0.02)
    tmp = tempfile.mktemp(suffix=".json")
    prof.export_chrome_trace(tmp)
    with open(tmp) as f:
        trace = json.load(f)
    os.unlink(tmp)
    ev = [e for e in trace["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
    ev.sort(key=lambda e: e["ts"])
    # split into steps at the largest idle gaps (the synchronize between steps)
    gaps = sorted(((ev[i + 1]["ts"] - (ev[i]["ts"] + ev[i]["dur"]), i) for i in range(len(ev) - 1)), reverse=True)[:2]
    cuts = sorted(i for _, i in gaps)
    steps = [ev[:cuts[0] + 1], ev[cuts[0] + 1:cuts[1] + 1], ev[cuts[1] + 1:]]
    st = steps[1]
    t0 = st[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in st)
    lines = [f"# Kernel timeline of one training step ({args.config}, {args.precision}, dropout {args.dropout}"
             f"{', receptive-field blocks' if args.receptive_field else ''}); "
             f"torch.profiler / CUPTI, step 2 of 3 after 4 warm-ups", "",
             f"step span {(t1 - t0) / 1e3:.3f} ms, {len(st)} device activities", "",
             "| start us | dur us | stream | idle before (any stream) us | activity |", "|---|---|---|---|---|"]
    busy_until = t0
    idle = 0.0
    for e in st:
        gap = max(0.0, e["ts"] - busy_until)
        idle += gap
        busy_until = max(busy_until, e["ts"] + e["dur"])
        lines.append(f"| {e['ts'] - t0:.0f} | {e['dur']:.0f} | {e.get('args', {}).get('stream', '?')} | {gap:.0f} | "
                     f"`{short(e['name'])}` |")
    lines += ["", f"device idle inside the step (no kernel on any stream): {idle / 1e3:.3f} ms of {(t1 - t0) / 1e3:.3f} ms"]
    agg = {}
    for e in st:
        a = agg.setdefault(short(e["name"]), [0, 0.0])
        a[0] += 1
        a[1] += e["dur"]
    lines += ["", "| kernel | launches | total us |", "|---|---|---|"]
    for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        lines.append(f"| `{n}` | {c} | {d:.0f} |")
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:6]))
    print(lines[-len(agg) - 5] if len(lines) > len(agg) + 5 else "")


if __name__ == "__main__":
    main()
