"""Generates ``tests/golden/*.npz`` by executing the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (``/root/reference`` must exist):

    python oracle/gen_golden.py

The reference's ``RelGATModel`` / ``RelGATLayer`` / scorers / losses / metric /
dataset code is imported verbatim through ``oracle/ref_shim.py`` (with the
``torch_scatter`` stand-in) and run on seeded inputs on the CPU.  Every array the
tests need — inputs, parameters, outputs, gradients — is stored, so the fixtures
can be checked on the GPU box where the reference is absent.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle.ref_shim import import_reference  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def make_graph(rng, n, e, r, *, isolated=0, self_loops=0, dup=0, hub=0, empty_rel=False):
    hi = n - isolated  # the last `isolated` nodes never appear as destinations
    src = rng.integers(0, n, size=e)
    dst = rng.integers(0, hi, size=e)
    rel = rng.integers(0, r - 1 if empty_rel else r, size=e)
    if self_loops:
        src[:self_loops] = dst[:self_loops]
    if dup:
        src[self_loops:self_loops + dup] = src[self_loops]
        dst[self_loops:self_loops + dup] = dst[self_loops]
        rel[self_loops:self_loops + dup] = rel[self_loops]
    if hub:
        dst[-hub:] = 3
    return src.astype(np.int64), dst.astype(np.int64), rel.astype(np.int64)


def model_case(name, *, n, e, r, d_in, f, h, layers, scorer, b, k, dtype, seed,
               projection=False, proj_layers=2, loss_type="margin", graph_kw=None,
               logit_scale=None, weights=(1.0, 1.0, 1.0, 0.0)):
    ref = import_reference()
    from relgat_projector.core.model.model import RelGATModel
    from relgat_projector.core.loss.relgat_loss import RelGATLoss
    from relgat_projector.core.loss.multi_objective_loss import MultiObjectiveRelLoss
    from relgat_projector.core.eval import RelgatEval

    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    src, dst, rel = make_graph(rng, n, e, r, **(graph_kw or {}))
    x0 = torch.from_numpy(rng.standard_normal((n, d_in))).to(dtype)
    edge_index = torch.from_numpy(np.stack([src, dst]))
    edge_type = torch.from_numpy(rel)
    model = RelGATModel(
        node_emb=x0, edge_index=edge_index, edge_type=edge_type, num_rel=r,
        scorer_type=scorer, gat_out_dim=f, gat_heads=h, dropout=0.0,
        relation_attn_dropout=0.0, gat_num_layers=layers,
        project_to_input_size=projection, projection_layers=proj_layers,
        projection_dropout=0.0, projection_hidden_dim=0,
    ).to(dtype)
    with torch.no_grad():  # non-trivial relation bias and (optionally) large logits
        gl = [model.gat_layer] if layers == 1 else list(model.gat_layers)
        for lyr in gl:
            lyr.rel_bias.copy_(torch.from_numpy(rng.standard_normal(r) * 0.1).to(dtype))
            if logit_scale is not None:
                for a in lyr.attn_vec:
                    a.mul_(logit_scale)
    model.train()

    pos_src = rng.integers(0, n, size=b)
    pos_dst = rng.integers(0, n, size=b)
    pos_rel = rng.integers(0, r, size=b)
    neg_dst = rng.integers(0, n, size=(k, b))
    src_ids = torch.from_numpy(np.concatenate([pos_src] + [pos_src] * k).astype(np.int64))
    rel_ids = torch.from_numpy(np.concatenate([pos_rel] + [pos_rel] * k).astype(np.int64))
    dst_ids = torch.from_numpy(np.concatenate([pos_dst] + [neg_dst[i] for i in range(k)]).astype(np.int64))

    rank = RelGATLoss(loss_type="self_adversarial_loss" if loss_type == "self_adv" else "margin",
                      self_adv_alpha=0.7, margin=1.0, clamp_limit=None, run_config={})
    out = {}
    if not projection:
        # trainer/relgat_projector.py:510-521 (no-projection branch)
        scores, _, dst_vec = model(src_ids, rel_ids, dst_ids, transform_to_input_if_possible=False)
        pos = scores[:b]
        neg = scores[b:].view(k, b).transpose(0, 1).contiguous()
        loss = rank.prepare_scores_and_compute_loss(pos_score=pos, neg_score=neg)
    else:
        # trainer/relgat_projector.py:587-655 (projection branch), same calls in the same order
        x = model.single_gat_step()
        ps, pd = x[src_ids[:b]], x[dst_ids[:b]]
        pos = model.scorer(ps, rel_ids[:b], pd)
        tr = model.scorer.transform(ps, rel_ids[:b])
        nd = x[dst_ids[b:]]
        neg = model.scorer(x[src_ids[b:]], rel_ids[b:], nd).view(b, k)
        ndv = nd.view(b, k, tr.shape[1]).permute(1, 0, 2).contiguous()
        pos = torch.nan_to_num(pos, nan=0.0, neginf=-1e9, posinf=1e9)
        neg = torch.nan_to_num(neg, nan=0.0, neginf=-1e9, posinf=1e9)
        multi = MultiObjectiveRelLoss(relgat_loss=rank, run_config={}, relgat_weight=weights[0],
                                      pos_cosine_weight=weights[1], neg_cosine_weight=weights[2],
                                      mse_weight=weights[3])
        loss = multi(pos_score=pos, neg_score=neg, transformed_src=tr, dst_vec=pd, neg_dst_vec=ndv)
        out["transformed_src"] = _np(tr)
    loss.backward()
    mrr, hits = RelgatEval.compute_mrr_hits(pos_score=pos.detach(), neg_score=neg.detach(),
                                            ks=tuple(range(1, k + 1)))
    with torch.no_grad():
        x_final = model.single_gat_step()
        layer0 = gl[0](x0, edge_index, edge_type)

    out.update(
        meta=np.array([n, e, r, d_in, f, h, layers, b, k, int(projection), proj_layers], dtype=np.int64),
        scorer=np.array(scorer), loss_type=np.array(loss_type), weights=np.array(weights, dtype=np.float64),
        x0=_np(x0), src=src, dst=dst, rel=rel,
        src_ids=_np(src_ids), rel_ids=_np(rel_ids), dst_ids=_np(dst_ids),
        layer0_out=_np(layer0), x_final=_np(x_final),
        pos=_np(pos), neg=_np(neg), loss=_np(loss), mrr=np.float64(mrr),
        hits=np.array([hits[i] for i in range(1, k + 1)], dtype=np.float64),
    )
    for name_p, p in model.state_dict().items():
        if name_p == "node_emb_fixed":
            continue
        out["param/" + name_p] = _np(p)
    for name_p, p in model.named_parameters():
        out["grad/" + name_p] = _np(p.grad) if p.grad is not None else np.zeros(0)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: loss={float(loss.detach()):.6f} mrr={mrr:.4f} size={os.path.getsize(path) / 1e3:.0f} kB")


def sampling_case():
    """Seed → shuffle/split → DataLoader batches with corrupted-tail negatives, exactly as the
    reference trainer consumes them (trainer:442-446; dataset/relgat_dataset.py:70-137)."""
    import_reference()
    from relgat_projector.utils.random_seed import RandomSeed
    from relgat_projector.dataset.relgat_dataset import RelGATDataset
    from relgat_projector.trainer.components.relgat_batching import concat_pos_negs_to_tensors

    n, t, r, d, k, bs = 200, 1500, 9, 4, 5, 64
    rng = np.random.default_rng(7)
    node2emb = {i: rng.standard_normal(d).astype(np.float32) for i in range(n)}
    rel2idx = {f"rel_{i}": i for i in range(r)}
    s = rng.integers(0, n, size=t)
    dd = rng.integers(0, n, size=t)
    rr = rng.integers(0, r, size=t)
    raw = [(int(a), int(b), f"rel_{int(c)}") for a, b, c in zip(s, dd, rr)]
    raw_copy = list(raw)
    RandomSeed(seed=1234, run_config={})
    ds = RelGATDataset(node2emb=node2emb, rel2idx=rel2idx, edge_index_raw=raw, train_ratio=0.9,
                       num_neg=k, train_batch_size=bs, eval_batch_size=bs, device="cpu", run_config={})
    out = dict(
        meta=np.array([n, t, r, d, k, bs, 1234], dtype=np.int64),
        raw_src=s.astype(np.int64), raw_dst=dd.astype(np.int64), raw_rel=rr.astype(np.int64),
        edge_index=_np(ds.edge_index), edge_type=_np(ds.edge_type),
        eval_edges=np.array([(a, b, rel2idx[c]) for a, b, c in ds.eval_edges], dtype=np.int64),
    )
    assert raw_copy != raw  # shuffled in place
    for bi, batch in enumerate(ds.train_loader):
        if bi >= 3:
            break
        pos, *negs = zip(*batch)
        a, b, c = concat_pos_negs_to_tensors(pos, negs, "cpu")
        out[f"batch{bi}_src"], out[f"batch{bi}_rel"], out[f"batch{bi}_dst"] = _np(a), _np(b), _np(c)
    path = os.path.join(GOLDEN, "sampling.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} size={os.path.getsize(path) / 1e3:.0f} kB")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    # (i) the tiny fp64 case of SURVEY.md §8(c)
    model_case("tiny_fp64", n=50, e=400, r=7, d_in=16, f=8, h=3, layers=2, scorer="distmult",
               b=16, k=3, dtype=torch.float64, seed=1)
    # (ii) head width of the named configs (F=200, H=4 → C=800), fp32, margin loss, one layer
    model_case("f200_fp32", n=120, e=900, r=11, d_in=64, f=200, h=4, layers=1, scorer="distmult",
               b=32, k=4, dtype=torch.float32, seed=2)
    # (iii) TransE + 3 layers + projection head + multi-objective loss (config 3's recipe)
    model_case("transe_proj_fp32", n=150, e=1200, r=9, d_in=48, f=24, h=8, layers=3, scorer="transe",
               b=32, k=4, dtype=torch.float32, seed=3, projection=True, proj_layers=2,
               weights=(1.0, 1.0, 1.0, 0.5))
    # (iv) adversarial graph: isolated destinations, self-loops, duplicates, a hub, an unused
    #      relation, logits of order ±80; single layer (gat_layer attribute), self-adversarial loss
    model_case("adversarial_fp32", n=300, e=4000, r=6, d_in=32, f=20, h=2, layers=1, scorer="distmult",
               b=24, k=2, dtype=torch.float32, seed=4, loss_type="self_adv", logit_scale=25.0,
               graph_kw=dict(isolated=40, self_loops=30, dup=25, hub=1500, empty_rel=True))
    # (v) K = 1, TransE without projection, fp64
    model_case("transe_fp64", n=60, e=500, r=5, d_in=12, f=12, h=2, layers=2, scorer="transe",
               b=20, k=1, dtype=torch.float64, seed=5)
    sampling_case()


if __name__ == "__main__":
    main()
