#!/usr/bin/env python
"""Benchmark of the RelGAT message-passing hot path (BASELINE.json metric: train edges/s, fwd+bwd).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c2] [--precision fp32]

One "step" = one full training step of the hot path over one batch: full-graph GAT forward
(L layers), fused gather-score of B*(1+K) triples, margin-ranking loss, backward to every
parameter gradient and the Adam update (reference trainer/relgat_projector.py:442-469).
``value`` = message-passing edges processed per second by the whole job = E / t_step.

Rank 0 prints ONE JSON line.  ``--impl reference`` times the UNMODIFIED reference (its own ``RelGATModel`` /
``RelGATLayer`` / scorer / loss classes, pip-installed into the git-ignored oracle/_ref by oracle/install_ref.py, with
the torch_scatter stand-in) on the host cores, on the SAME configuration as this arm (kind "reference"; the number of
timed steps is capped by a wall-clock budget and reported); without that install it falls back to the oracle port on a
scaled-down sample (kind "port").  No product code runs in that arm.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "relgat_train_edges_per_sec_fwd_bwd"
UNIT = "edges/s"
CPU_SAMPLE_SCALE = 20    # oracle-port fallback: the named config at 1/20 of the nodes and edges
CPU_BASELINE_SCALE = 5   # cpu_baseline leg of this arm: the reference at 1/5 scale (~10-20 s of CPU work)
REF_BUDGET_S = float(os.environ.get("RELGAT_REF_BUDGET_S", "240"))  # wall-clock budget of the reference arm's timed steps


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), tflops=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json: copy bandwidth for HBM, sustained bf16 GEMM figure for tensor)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# per-kernel event timing (monkeypatches the ops wrappers for a few extra steps)
# ---------------------------------------------------------------------------------------------
class KernelProfiler:
    NAMES = ["split_bf16", "gemm", "edge_fwd", "edge_bwd_prep", "edge_bwd_src", "edge_bwd_rel", "score_fwd",
             "score_bwd", "index_add_sorted", "margin_loss", "rank_loss", "recon_loss", "zero_rows", "pull_rows",
             "edge_bwd_beta", "gemm_dx_prep"]

    def __init__(self):
        from relgat_projector_b200 import ops
        self.ops = ops
        self.records = []  # (name, tag, start, end)
        self.orig = {}
        self.main = torch.cuda.current_stream()

    def __enter__(self):
        for name in self.NAMES:
            fn = getattr(self.ops, name)
            self.orig[name] = fn
            setattr(self.ops, name, self._wrap(name, fn))
        return self

    def _wrap(self, name, fn):
        def wrapped(*a, **kw):
            tag = name
            if name == "gemm":
                tag = f"gemm[M={a[4]},N={a[5]},K={a[6]},{'mn' if a[1] else 'k'}{'mn' if a[3] else 'k'}]"
            if name == "gemm_dx_prep":
                tag = f"gemm[M={a[2]},N={a[3]},K={a[4]},kk+prep]"
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            side = torch.cuda.current_stream() != self.main
            s.record()
            out = fn(*a, **kw)
            e.record()
            self.records.append((name, tag + (" (side stream)" if side else ""), s, e))
            return out
        return wrapped

    def __exit__(self, *exc):
        for name, fn in self.orig.items():
            setattr(self.ops, name, fn)

    def table(self, steps: int):
        torch.cuda.synchronize()
        agg = {}
        for name, tag, s, e in self.records:
            d = agg.setdefault(tag, {"kernel": name, "calls": 0, "ms": 0.0})
            d["calls"] += 1
            d["ms"] += s.elapsed_time(e)
        for d in agg.values():
            d["ms_per_step"] = d["ms"] / steps
            d["avg_ms"] = d["ms"] / d["calls"]
        return agg


# ---------------------------------------------------------------------------------------------
# algorithmic bytes / flops (SURVEY.md §8(d) conventions; stated in DESIGN.md)
# ---------------------------------------------------------------------------------------------
def kernel_work(cfg, E, n_chunks, precision, use_ds=True):
    N, H, F, R = cfg["N"], cfg["H"], cfg["F"], cfg["R"]
    C = H * F
    s = 4
    sf = 4 if precision == "fp32" else 2       # bytes per stored feature element (P and G rows)
    plane_b = 4 if precision == "fp32" else 2  # bytes per element of the bf16 (hi[, lo]) planes
    w = {
        # gathers P[src] + (src, rel) ids; writes out (fp32), z, (max, 1/den) and bias; reads rowptr
        "edge_fwd": E * (C * sf + 8) + N * (C * s + 4 + 4 + H * 8) + E * H * 4,
        "edge_fwd_act": N * C * plane_b,
        # own P row + gather G[dst] + (slot, dst, rel) ids + z, (max, 1/den), t; writes dP planes and dz
        "edge_bwd_src": (N * C * sf + E * (C * sf + 12 + 4 * H * 4) + N * C * plane_b
                         + (N * H * R * plane_b if use_ds else E * H * 4)),  # dS columns, or dz per edge
        # gathers P[src] + (slot, src, dst) ids + dz + hsum; writes chunk partials
        "edge_bwd_rel": E * (C * sf + 12 + 2 * H * 4) + n_chunks * C * s * 2,
    }
    # bwd_prep, average over the L calls of a step.  Hidden layers: read dY and out, write G (ELU').  Last layer,
    # fp32: G aliases dY and only the <= 2*B*(1+K) rows the loss touched are read (t / hsum of the rest: memset);
    # bf16 storage: dense read of dY and out, bf16 G written.
    L = cfg["L"]
    rows = min(N, 2 * cfg["B"] * (1 + cfg["K"]))
    if precision == "fp32":
        total = (L - 1) * (3 * N * C * s) + 2 * rows * C * s + L * 2 * N * H * 4
    else:
        total = L * (2 * N * C * s + N * C * sf + 2 * N * H * 4)
    w["edge_bwd_prep"] = total // L
    return w


def step_bytes_survey(cfg, E, precision):
    """bytes_step of SURVEY.md §8(d): sum over layers of FWD_l + BWD_l (s = 4 bytes)."""
    N, H, F, L, D = cfg["N"], cfg["H"], cfg["F"], cfg["L"], cfg["D_in"]
    C, s, tot = H * F, 4, 0
    for l in range(L):
        d_in = D if l == 0 else C
        fwd = N * d_in * s + N * C * s + E * (C * s + 8) + N * (C * s + 4)
        bwd = E * (2 * C * s + 16 + 8 * H) + N * 2 * C * s + N * (d_in * s + C * s)
        if l > 0:
            bwd += N * (C * s + d_in * s)
        tot += fwd + bwd
    return tot


# ---------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_port_setup(cfg, scale, seed=42):
    from relgat_projector_b200 import synthetic as S
    from relgat_projector_b200.layer import RelGATLayer
    n, t = max(cfg["N"] // scale, 64), max(cfg["T"] // scale, 256)
    kg = S.tensor_kg(n, t, cfg["R"], cfg["D_in"], seed=seed, device="cpu")
    torch.manual_seed(seed)
    H, F, L = cfg["H"], cfg["F"], cfg["L"]
    layers, in_dim = [], cfg["D_in"]
    for _ in range(L):
        lyr = RelGATLayer(in_dim, F, cfg["R"], heads=H, dropout=0.0)  # same init as the product / reference
        layers.append({"W": [p.weight for p in lyr.proj], "A": list(lyr.attn_vec), "beta": lyr.rel_bias})
        in_dim = H * F
    rel_emb = torch.nn.Parameter(torch.empty(cfg["R"], H * F))
    torch.nn.init.xavier_uniform_(rel_emb)
    params = [p for lp in layers for p in (*lp["W"], *lp["A"], lp["beta"])] + [rel_emb]
    opt = torch.optim.Adam(params, lr=2e-4)
    g = torch.Generator().manual_seed(seed)
    return kg, layers, rel_emb, params, opt, g


def cpu_port_step(cfg, kg, layers, rel_emb, opt, g):
    from oracle import relgat_oracle as O  # bench.py's cpu_baseline / --impl reference legs only
    from relgat_projector_b200 import synthetic as S
    b, k = cfg["B"], cfg["K"]
    src, rel, dst = S.sample_batch(kg.train_triples, kg.node_emb.size(0), b, k, g)
    opt.zero_grad(set_to_none=True)
    loss, _, _ = O.train_step_port(kg.node_emb, layers, rel_emb, kg.edge_index, kg.edge_type, src, rel, dst,
                                   scorer=cfg["scorer"], b=b, k=k, margin=1.0)
    loss.backward()
    opt.step()
    return float(loss.detach())


def run_cpu_port(cfg, steps, warmup, scale):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kg, layers, rel_emb, params, opt, g = cpu_port_setup(cfg, scale)
    E = int(kg.edge_index.size(1))
    for _ in range(warmup):
        cpu_port_step(cfg, kg, layers, rel_emb, opt, g)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port_step(cfg, kg, layers, rel_emb, opt, g)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    sample = (f"{cfg['name']} at 1/{scale} scale: N={kg.node_emb.size(0)} nodes, E={E} edges, R={cfg['R']}, "
              f"D_in={cfg['D_in']}, L={cfg['L']}, H={cfg['H']}, F={cfg['F']}, B={cfg['B']}, K={cfg['K']}; "
              f"{warmup} warm-up + {steps} timed steps, fp32, torch {torch.__version__} CPU")
    return dict(value=E / dt, unit=UNIT, cores=cores, kind="port", sample=sample, ms_per_step=dt * 1e3, edges=E)



# ---------------------------------------------------------------------------------------------
# reference arm: the UNMODIFIED reference on the host cores (no product code)
# ---------------------------------------------------------------------------------------------
def _standalone_synthetic():
    """The seeded KG generator, loaded as a plain file (no import of the product package: this arm must not depend on
    it).  Same seeds -> the same graph and batches as the GPU arm."""
    import importlib.util
    name = "_relgat_bench_synthetic"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "relgat_projector_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_installed() -> bool:
    from oracle import ref_shim
    return ref_shim.reference_available()


def run_cpu_reference(cfg, steps, scale=1, budget_s=REF_BUDGET_S, seed=42):
    """Training steps of the reference's own classes (trainer/relgat_projector.py:442-469, 498-557) on the CPU."""
    from oracle import ref_shim
    ref_shim.import_reference()
    from relgat_projector.core.loss.multi_objective_loss import MultiObjectiveRelLoss
    from relgat_projector.core.loss.relgat_loss import RelGATLoss
    from relgat_projector.core.model.model import RelGATModel
    S = _standalone_synthetic()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, t = max(cfg["N"] // scale, 64), max(cfg["T"] // scale, 256)
    kg = S.tensor_kg(n, t, cfg["R"], cfg["D_in"], seed=seed, device="cpu")
    E = int(kg.edge_index.size(1))
    torch.manual_seed(seed)
    model = RelGATModel(node_emb=kg.node_emb, edge_index=kg.edge_index, edge_type=kg.edge_type, num_rel=cfg["R"],
                        scorer_type=cfg["scorer"], gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=0.0,
                        relation_attn_dropout=0.0, gat_num_layers=cfg["L"], project_to_input_size=cfg["proj"],
                        projection_layers=2)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=2e-4)
    rank = RelGATLoss("margin_ranking_loss", None, 1.0, 20, {})
    multi = MultiObjectiveRelLoss(relgat_loss=rank, run_config={}) if cfg["proj"] else None
    b, k = cfg["B"], cfg["K"]
    gen = torch.Generator().manual_seed(seed)

    def step():
        src, rel, dst = S.sample_batch(kg.train_triples, n, b, k, gen)
        opt.zero_grad(set_to_none=True)
        if not cfg["proj"]:  # trainer:510-521 + _split_scores :657-676
            scores, _, _ = model(src, rel, dst, transform_to_input_if_possible=False)
            loss = rank.prepare_scores_and_compute_loss(pos_score=scores[:b],
                                                        neg_score=scores[b:].view(k, b).transpose(0, 1).contiguous())
        else:  # trainer:587-655
            x = model.single_gat_step()
            ps, pd = x[src[:b]], x[dst[:b]]
            pos = model.scorer(ps, rel[:b], pd)
            tr = model.scorer.transform(ps, rel[:b])
            nd = x[dst[b:]]
            neg = model.scorer(x[src[b:]], rel[b:], nd).view(b, k)
            loss = multi(pos_score=pos, neg_score=neg, transformed_src=tr, dst_vec=pd,
                         neg_dst_vec=nd.view(b, k, tr.shape[1]).permute(1, 0, 2).contiguous())
        loss.backward()
        opt.step()
        return float(loss.detach())

    t0 = time.perf_counter()
    step()  # warm-up (allocator, thread pools)
    t_warm = time.perf_counter() - t0
    n_timed = max(1, min(int(steps), int(budget_s / max(t_warm, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(n_timed):
        step()
    dt = (time.perf_counter() - t0) / n_timed
    what = "the full configuration" if scale == 1 else f"1/{scale} of the nodes and triplets"
    sample = (f"{cfg['name']}, {what}: N={n} nodes, E={E} message-passing edges, R={cfg['R']}, D_in={cfg['D_in']}, "
              f"L={cfg['L']}, H={cfg['H']}, F={cfg['F']}, B={b}, K={k}; UNMODIFIED reference classes (RelGATModel, "
              f"RelGATLoss; torch_scatter stand-in) on {cores} host threads; 1 warm-up + {n_timed} timed steps "
              f"({steps} requested, capped by a {budget_s:.0f} s budget), fp32, torch {torch.__version__} CPU")
    return dict(value=E / dt, unit=UNIT, cores=cores, kind="reference", sample=sample, ms_per_step=dt * 1e3, edges=E,
                steps=n_timed, scale=scale)


def workload_config(cfg_name, cfg, E):
    """``config`` of the JSON line: identical in both arms."""
    return {"workload": f"{cfg_name}: synthetic KG {cfg['N']} nodes / {cfg['T']} triplets ({E} message-passing edges) / "
                        f"{cfg['R']} relations, {cfg['D_in']}-d, {cfg['L']} layers, {cfg['H']} heads, gat-out-dim "
                        f"{cfg['F']}, {cfg['scorer']}, batch {cfg['B']}, num-neg {cfg['K']}",
            "precision": "fp32",
            "step": "full-graph GAT fwd + gather-score + margin loss + bwd + Adam",
            "l2": "inputs_exceed_L2 (P and G are ~1 GB each vs 126 MB L2)"}


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, help="c1|c2|c3|c4|tiny (default c2: on N > 1 GPUs every rank holds one c2-sized share of an N-times larger graph = "
                                                         "weak scaling; an explicitly named config is sharded as it is = strong scaling)")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--exchange", default="peer", choices=["peer", "halo", "allgather"],
                    help="multi-GPU only: peer tables read over NVLink (default), or an NCCL exchange of halo / all rows")
    ap.add_argument("--halo", default="bf16", choices=["bf16", "fp32"],
                    help="multi-GPU, peer tables: halo rows cross NVLink rounded to bf16 (default; half the link bytes, "
                         "stated tolerance 2e-2) or as fp32 (results equal to one GPU up to summation order)")
    ap.add_argument("--locality", type=float, default=0.0,
                    help="multi-GPU supplementary runs: fraction of edges whose head is drawn from the tail's node block "
                         "(default 0: uniformly random graph, the headline workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt-precision", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=3)
    ap.add_argument("--backward", default="dense", choices=["dense", "sparse"],
                    help="dense (headline): the by-source pass gathers G[dst] for every edge, SURVEY.md §8(d)'s unit of work; "
                         "sparse: the backward covers only the rows whose gradient can be non-zero (the library's default; "
                         "same gradients) — always reported beside the headline as exact_sparse_backward")
    ap.add_argument("--ref-scale", type=int, default=1,
                    help="--impl reference: run the configuration at 1/SCALE of its nodes and triplets (default 1: in full)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from relgat_projector_b200 import synthetic as S
    cfg_name = args.config or "c2"
    cfg = dict(S.CONFIGS[cfg_name], name=cfg_name)

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the per-GPU workload of this arm (weak scaling: every rank holds one such graph); one CPU host runs it once
        if reference_installed():
            r = run_cpu_reference(cfg, args.steps, scale=args.ref_scale)
            n_warm = 1
        else:
            scale = CPU_SAMPLE_SCALE if cfg["N"] >= 100_000 else 1
            n_warm = max(args.warmup, 1)
            r = dict(run_cpu_port(cfg, args.steps, n_warm, scale), steps=args.steps, scale=scale)
        e_full = int(0.9 * cfg["T"])
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": n_warm, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": workload_config(cfg_name, cfg, e_full if r["scale"] == 1 else r["edges"]),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "steps_requested": args.steps,
        }
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    if world > 1:
        from relgat_projector_b200 import dist as RD  # destination-range partitioned path
        cpu_fn = None
        if not args.no_cpu_baseline and reference_installed():  # rank 0, bounded sample of the per-GPU workload
            cpu_fn = lambda: run_cpu_reference(cfg, steps=1, scale=CPU_BASELINE_SCALE if cfg["N"] >= 100_000 else 1,  # noqa: E731
                                               budget_s=30.0)
        return RD.bench_main(args, cfg, rank, world, local_rank, METRIC, UNIT, load_peaks, ClockSampler, cpu_fn=cpu_fn)

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    import relgat_projector_b200 as R
    from relgat_projector_b200 import functional as RFn, loss as L, ops
    RFn.SPARSE_BWD = args.backward == "sparse"

    kg = S.tensor_kg(cfg["N"], cfg["T"], cfg["R"], cfg["D_in"], seed=42, device=str(dev))
    E = int(kg.edge_index.size(1))
    torch.manual_seed(42)
    model = R.RelGATModel(kg.node_emb, kg.edge_index, kg.edge_type, num_rel=cfg["R"], scorer_type=cfg["scorer"],
                          gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=0.0, gat_num_layers=cfg["L"],
                          project_to_input_size=cfg["proj"], projection_layers=2, precision=args.precision).to(dev)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, fused=True)  # the reference trainer's Adam (trainer:265); torch's single-kernel implementation
    rank_loss = L.RelGATLoss("margin", None, 1.0, None, {})
    b, k = cfg["B"], cfg["K"]
    gen = torch.Generator().manual_seed(42)
    n_pool = 8
    host_batches = [tuple(t.pin_memory() for t in S.sample_batch(kg.train_triples.cpu(), cfg["N"], b, k, gen))
                    for _ in range(n_pool)]
    dev_batches = [tuple(t.to(dev) for t in hb) for hb in host_batches]

    multi_loss = L.MultiObjectiveRelLoss(relgat_loss=rank_loss, run_config={}) if cfg["proj"] else None

    def train_step(src, rel, dst):
        # the trainer's step (reference trainer/relgat_projector.py:442-469): _calculate_loss, backward, Adam
        opt.zero_grad(set_to_none=True)
        _, _, loss, *_ = L.calculate_loss(model, src, rel, dst, b, rank_loss, multi_loss)
        loss.backward()
        opt.step()
        return loss

    step_stats = {}

    def timed(fn, steps, tag=None):
        """Mean step time over EXACTLY ``steps`` steps (one pair of CUDA events around the whole region); the events
        recorded after every step only feed the per-step spread reported beside it."""
        torch.cuda.synchronize()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        marks[0].record()
        for i in range(steps):
            fn(i)
            marks[i + 1].record()
        torch.cuda.synchronize()
        if tag and steps:
            per = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(steps))
            step_stats[tag] = {"min_ms": round(per[0], 3), "median_ms": round(per[len(per) // 2], 3),
                               "max_ms": round(per[-1], 3)}
        return marks[0].elapsed_time(marks[-1]) / max(steps, 1)

    warm = max(args.warmup, 3)
    for i in range(warm):
        train_step(*dev_batches[i % n_pool])
    ops.LAUNCHES = 0
    with ClockSampler(local_rank) as clocks:
        ms = timed(lambda i: train_step(*dev_batches[i % n_pool]), args.steps, "value")
    launches = ops.LAUNCHES

    # end to end through the public API: ids come from pinned host memory, the loss goes back
    def e2e_step(i):
        hb = host_batches[i % n_pool]
        src, rel, dst = (t.to(dev, non_blocking=True) for t in hb)
        return float(train_step(src, rel, dst).item())
    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps, "e2e")
    h2d = 3 * b * (1 + k) * 8

    # per-kernel device times (extra steps, CUDA events around every C-ABI call on the launch stream)
    psteps = max(args.profile_steps, 1)
    with KernelProfiler() as prof:
        for i in range(psteps):
            train_step(*dev_batches[i % n_pool])
        table = prof.table(psteps)
    g = model._graph()
    work = kernel_work(cfg, E, g.n_chunks, args.precision, use_ds=RFn.USE_DS)
    peaks = load_peaks()
    kernels = {}
    for tag, d in sorted(table.items(), key=lambda kv: -kv[1]["ms_per_step"]):
        ent = {"calls_per_step": d["calls"] / psteps, "avg_ms": round(d["avg_ms"], 4),
               "ms_per_step": round(d["ms_per_step"], 4), "share_of_step": round(d["ms_per_step"] / ms, 4)}
        if args.backward == "sparse" and d["kernel"] in ("edge_bwd_src", "edge_bwd_prep"):
            # rows restricted to the batch's in-neighbourhood: §8(d)'s dense byte count does not describe this launch
            ent["note"] = "restricted to the rows that can be non-zero; no algorithmic-byte figure (see profiles/r02_ncu_sparse_bwd_src.md for its measured DRAM traffic)"
        elif d["kernel"] in work:
            bytes_ = work[d["kernel"]] + (work["edge_fwd_act"] * (cfg["L"] - 1) / cfg["L"] if d["kernel"] == "edge_fwd" else 0)
            ent["algorithmic_bytes"] = int(bytes_)
            ent["gbs"] = round(bytes_ / (d["avg_ms"] * 1e-3) / 1e9, 1)
            ent["frac_hbm"] = round(ent["gbs"] / peaks["hbm_gbs"], 4)
        elif d["kernel"] in ("gemm", "gemm_dx_prep"):
            dims = dict(kv.split("=") for kv in tag[tag.index("[") + 1:tag.rindex(",")].split(","))
            flops = 2.0 * int(dims["M"]) * int(dims["N"]) * int(dims["K"]) * (3 if args.precision == "fp32" else 1)
            ent["tensor_flops"] = flops
            ent["tflops"] = round(flops / (d["avg_ms"] * 1e-3) / 1e12, 1)
            ent["frac_tensor"] = round(ent["tflops"] / peaks["tflops"], 4)
        if tag.endswith("(side stream)"):
            # parameter-only operand prep (fold_operands) launched beside the layer's GEMM: its elapsed time is the
            # big kernel's, not its own (isolated: 8-30 us per launch, profiles/r02_bench_c2_fp32_src3.json)
            ent["overlapped"] = True
            ent.pop("tflops", None); ent.pop("frac_tensor", None)
        if not RFn.USE_DS and (d["kernel"] == "edge_bwd_rel" or (d["kernel"] == "gemm" and tag.endswith("mnmn]"))):
            # the by-relation pass runs on a side stream beside the dW GEMM: both elapsed times include
            # the other's interference (isolated: 0.69 ms at 96 % of HBM peak; 1.14 ms) — not "dominant"
            ent["overlapped"] = True
        kernels[tag] = ent
    top_tag = next(t for t, e in kernels.items() if not e.get("overlapped"))
    top = kernels[top_tag]
    if "tflops" in top:
        roofline = {"kernel": top_tag, "bound": "tensor", "achieved": top["tflops"], "peak": peaks["tflops"],
                    "unit": "TFLOP/s", "frac": top["frac_tensor"], "traffic": None}
    else:
        roofline = {"kernel": top_tag, "bound": "hbm", "achieved": top.get("gbs"), "peak": peaks["hbm_gbs"],
                    "unit": "GB/s", "frac": top.get("frac_hbm"), "traffic": None}
    roofline["peak_source"] = peaks["source"]
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of this same command
    # (profiles/r02_ncu_traffic.json; a number taken under a profiler in an earlier call, labelled as such — ncu cannot
    # run inside a timed benchmark); only meaningful for the configuration it was taken on
    # (the edge kernels: the capture of the same launches at the same shapes by tools/profile_layer.py)
    src3 = RFn.SRC_V3 and RFn.USE_DS
    sub = {"edge_fwd": "edge_fwd_kernel", "edge_bwd_src": "bwd_src3_kernel" if src3 else "bwd_src_kernel",
           "edge_bwd_rel": "bwd_rel_kernel", "edge_bwd_prep": "bwd_prep_kernel",
           "gemm": "gemm_bf16_tcgen05"}.get(top_tag.split("[")[0])
    for tfile in ("r02_ncu_src3_traffic.json", "r02_ncu_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tfile)
        if roofline["traffic"] is not None or not (sub and cfg_name == "c2" and args.precision == "fp32" and os.path.exists(tpath)):
            continue
        with open(tpath) as f:
            for kname, rec in json.load(f).items():
                if sub in kname:
                    v = rec["dram_bytes_per_launch"]
                    roofline["traffic"] = int(sum(v) / len(v))
                    roofline["traffic_source"] = (f"profiles/{tfile}: dram__bytes_read.sum + dram__bytes_write.sum of "
                                                  "this kernel in the committed ncu --set full capture of the same launch "
                                                  "(not measured in this run)")
                    break
    sbytes = step_bytes_survey(cfg, E, args.precision)
    step_roof = {"algorithmic_bytes_per_step": sbytes, "achieved_gbs": round(sbytes / (ms * 1e-3) / 1e9, 1),
                 "frac_of_hbm_peak": round(sbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4),
                 "formula": "SURVEY.md §8(d) bytes_step, fp32"}

    # side record: the same step with the backward restricted to the rows of the output gradient that can be non-zero
    # (the library default, functional.SPARSE_BWD / COMPACT_BWD): same gradients, a different unit of work than §8(d)'s
    sparse_rec = None
    if args.backward == "dense" and not args.no_alt_precision and RFn.USE_DS:
        RFn.SPARSE_BWD = True
        for i in range(3):
            train_step(*dev_batches[i % n_pool])
        ms_sp = timed(lambda i: train_step(*dev_batches[i % n_pool]), max(args.steps // 2, 5))
        with KernelProfiler() as prof:
            for i in range(psteps):
                train_step(*dev_batches[i % n_pool])
            tsp = prof.table(psteps)
        RFn.SPARSE_BWD = False
        sparse_rec = {"ms_per_step": ms_sp, "value": E / (ms_sp * 1e-3), "unit": UNIT,
                      "speedup_vs_dense": round(ms / ms_sp, 4),
                      "edge_bwd_src_avg_ms": [round(d["avg_ms"], 4) for t_, d in tsp.items() if d["kernel"] == "edge_bwd_src"],
                      "compacted": bool(RFn.COMPACT_BWD),
                      "note": "full-graph forward; backward restricted to the rows of dL/d out that can be non-zero "
                              "(the batch rows at the last layer, their in-neighbourhood one layer down): edges into "
                              "exact-zero rows are skipped (bit-identical gradients) and, compacted, the dP rows and the "
                              "weight-gradient / dX GEMMs cover those rows only (gradients equal to rounding, 2e-5; "
                              "tests/test_gpu_fused.py); edges/s still counts all E edges of the graph"}

    # side record: the same step on the batch's receptive-field blocks — per step, the L-hop
    # in-neighbourhood of the batch's nodes is extracted on the GPU (inside the timed region) and the same kernels run on
    # it; same loss and gradients (tests/test_gpu_fused.py), a DIFFERENT unit of work: only `block_edges_per_step` of the
    # L*E layer-edges are processed, so its edges/s is not comparable with the headline's and is labelled as such
    rf_rec = None
    if not args.no_alt_precision:
        model.receptive_field = True
        for i in range(3):
            train_step(*dev_batches[i % n_pool])
        ms_rf = timed(lambda i: train_step(*dev_batches[i % n_pool]), max(args.steps // 2, 5))
        rf_rec = {"ms_per_step": ms_rf, "steps_per_sec": 1e3 / ms_rf, "speedup_vs_full_graph_step": round(ms / ms_rf, 3),
                  "block_edges_per_step": int(model.last_block_edges), "layer_edges_full_graph": cfg["L"] * E,
                  "graph_edges_per_sec_equivalent": E / (ms_rf * 1e-3),
                  "mode": getattr(model, "receptive_field_mode", None),
                  "note": "receptive-field pruning: row-set plan + forward + loss + backward + Adam per step; "
                          "'graph_edges_per_sec_equivalent' divides ALL E edges by the step time for comparison only"}
        model.receptive_field_mode = "blocks"  # the alternative implementation: per-batch bipartite sub-indexes
        for i in range(3):
            train_step(*dev_batches[i % n_pool])
        rf_rec["blocks_mode_ms_per_step"] = timed(lambda i: train_step(*dev_batches[i % n_pool]), max(args.steps // 2, 5))
        model.receptive_field_mode = "masked"
        model.receptive_field = False

    # secondary record: single-pass bf16 tensor-core operands (stated tolerance 2e-2, tests/test_gpu_model.py)
    alt = None
    if args.precision == "fp32" and not args.no_alt_precision:
        model = opt = None
        torch.cuda.empty_cache()
        torch.manual_seed(42)
        model = R.RelGATModel(kg.node_emb, kg.edge_index, kg.edge_type, num_rel=cfg["R"], scorer_type=cfg["scorer"],
                              gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=0.0, gat_num_layers=cfg["L"],
                              project_to_input_size=cfg["proj"], projection_layers=2, precision="bf16").to(dev)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=2e-4, fused=True)  # the reference trainer's Adam (trainer:265); torch's single-kernel implementation
        for i in range(3):
            train_step(*dev_batches[i % n_pool])
        ms_alt = timed(lambda i: train_step(*dev_batches[i % n_pool]), max(args.steps // 2, 5))
        alt = {"dtype": "bf16 feature storage + single-pass bf16 tensor-core operands, fp32 accumulate", "value": E / (ms_alt * 1e-3),
               "unit": UNIT, "ms_per_step": ms_alt, "tolerance": "2e-2 relative (stated, tested)"}

    # side record: the same step with the reference's training dropout (training_scripts/run-relgat-trainer-base-model.sh:
    # dropout 0.3) — masks drawn per step, applied inside the fused edge kernels
    drop_rec = None
    if not args.no_alt_precision:
        model = opt = None
        torch.cuda.empty_cache()
        torch.manual_seed(42)
        model = R.RelGATModel(kg.node_emb, kg.edge_index, kg.edge_type, num_rel=cfg["R"], scorer_type=cfg["scorer"],
                              gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=0.3, gat_num_layers=cfg["L"],
                              project_to_input_size=cfg["proj"], projection_layers=2, precision=args.precision).to(dev)
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=2e-4, fused=True)  # the reference trainer's Adam (trainer:265); torch's single-kernel implementation
        for i in range(3):
            train_step(*dev_batches[i % n_pool])
        ms_drop = timed(lambda i: train_step(*dev_batches[i % n_pool]), max(args.steps // 2, 5))
        drop_rec = {"dropout": 0.3, "relation_attn_dropout": 0.0, "ms_per_step": ms_drop, "value": E / (ms_drop * 1e-3),
                    "unit": UNIT, "slowdown_vs_dropout0": round(ms_drop / ms - 1.0, 4),
                    "note": "one fused autograd node as with dropout 0; keep-bit masks from Philox, applied in edge_fwd / bwd_prep"}

    cpu = None
    if not args.no_cpu_baseline:
        big = cfg["N"] >= 100_000
        if reference_installed():  # a bounded sample (~10-30 s of CPU work); the full configuration is --impl reference
            r = run_cpu_reference(cfg, steps=1, scale=CPU_BASELINE_SCALE if big else 1, budget_s=30.0)
        else:
            r = run_cpu_port(cfg, steps=2, warmup=1, scale=CPU_SAMPLE_SCALE if big else 1)
        cpu = {kk: r[kk] for kk in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": E / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32" if args.precision == "fp32" else "bf16", "data": "synthetic",
        "config": dict(workload_config(cfg_name, cfg, E), precision=args.precision),
        "backward": args.backward,
        "precision_note": ("fp32 storage; tensor-core GEMMs on bf16 hi/lo splits (3 passes, ~fp32 accuracy)"
                           if args.precision == "fp32" else
                           "bf16 storage of P / G / dP rows; single-pass bf16 tensor-core GEMMs; fp32 accumulate"),
        "layer_edges_per_sec": cfg["L"] * E / (ms * 1e-3), "per_step_spread": step_stats,
        "clocks": clocks.summary(),
        "e2e": {"value": E / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "optimizer": "torch.optim.Adam(lr=2e-4, fused=True): the reference trainer's optimizer (trainer/relgat_projector.py:265) in torch's single-kernel implementation",
        "roofline": roofline, "step_roofline": step_roof, "kernels": kernels, "cpu_baseline": cpu,
        "alt_precision": alt, "training_dropout": drop_rec, "exact_sparse_backward": sparse_rec, "receptive_field": rf_rec,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
