"""Trainer-level drop-in proof: the reference's OWN ``RelGATTrainer`` runs twice on the same
seeded knowledge graph — once exactly as shipped on the CPU (its torch + torch_scatter path), once on the GPU after
the four class assignments of INTEGRATION.md §1 (optionally plus the trainer-level seam of §1b) — and the loss curve of
a short epoch and the MRR / Hits@k of ``evaluate()`` are compared.

The reference package is the UNMODIFIED one that ``oracle/install_ref.py`` pip-installed into the git-ignored
``oracle/_ref`` (``/root/reference`` does not exist on the GPU box); ``torch_scatter`` is the stand-in of
``oracle/standin``.  The test skips when that install is absent.
"""
import contextlib
import io
import random

import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.gpu

# c1's layer shapes (1024-d inputs, 2 layers, 4 heads x 200) on a graph small enough for the CPU run of the
# reference trainer to take seconds; num_neg = 10 so that Hits@10 exists (the metric ranks against the sampled negatives)
N_NODES, N_TRIPLES, N_REL, D_IN, HEADS, F_OUT, LAYERS, BATCH, NUM_NEG = 1500, 4000, 12, 1024, 4, 200, 2, 256, 10
KS = (1, 3, 10)
LOSS_TOL = 2e-3   # relative, per step: fp32 GEMM-order differences amplified by a few Adam steps
METRIC_TOL = 2e-2  # absolute on MRR / Hits@k (a handful of near-ties out of 400 eval triples may flip)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not ref_shim.reference_available():
        pytest.skip("reference package not installed (python oracle/install_ref.py in the build container)")
    return "cuda:0"


def _inputs(seed=7):
    from relgat_projector_b200 import synthetic as S
    return S.reference_inputs(N_NODES, N_TRIPLES, N_REL, D_IN, seed=seed)


def _run_trainer(device, tmp_path, *, projection, self_adv, patch, fused_seam=False):
    """One epoch of the reference trainer + evaluate().  ``patch``: apply INTEGRATION.md §1's assignments."""
    ref_shim.import_reference()
    import relgat_projector.core.model.model as ref_model_mod
    import relgat_projector.trainer.relgat_projector as ref_trainer_mod
    import relgat_projector_b200 as b200

    saved = {(m, n): getattr(m, n) for m, n in
             ((ref_model_mod, "RelGATLayer"), (ref_model_mod, "DistMultScorer"), (ref_model_mod, "TransEScorer"),
              (ref_model_mod, "ProjectionHead"), (ref_trainer_mod, "RelGATModel"), (ref_trainer_mod, "RelGATLoss"),
              (ref_trainer_mod, "MultiObjectiveRelLoss"))}
    saved_calc = ref_trainer_mod.RelGATTrainer._calculate_loss
    try:
        if patch:
            ref_model_mod.RelGATLayer = b200.RelGATLayer          # INTEGRATION.md §1, the four assignments
            ref_model_mod.DistMultScorer = b200.DistMultScorer
            ref_model_mod.TransEScorer = b200.TransEScorer
            ref_trainer_mod.RelGATModel = b200.RelGATModel
            if fused_seam:                                        # §1b: loss classes + the trainer-level seam
                ref_trainer_mod.RelGATLoss = b200.loss.RelGATLoss
                ref_trainer_mod.MultiObjectiveRelLoss = b200.loss.MultiObjectiveRelLoss

                def _calc(self, src_ids, rel_ids, dst_ids, pos_examples_in_batch, phase):
                    pos, neg, loss, mse, cpos, cneg = b200.loss.calculate_loss(
                        self.model, src_ids, rel_ids, dst_ids, pos_examples_in_batch, self.relgat_loss, self.multi_loss)
                    item = lambda v: None if v is None else float(v)  # noqa: E731
                    return pos, neg, loss, item(mse), item(cpos), item(cneg)

                ref_trainer_mod.RelGATTrainer._calculate_loss = _calc
        node2emb, rel2idx, raw = _inputs()
        with contextlib.redirect_stdout(io.StringIO()):
            tr = ref_trainer_mod.RelGATTrainer(
                run_config={"early_stop_patience": 1000}, node2emb=node2emb, rel2idx=rel2idx, edge_index_raw=list(raw),
                train_batch_size=BATCH, eval_batch_size=BATCH, device=device, seed=123, lr=2e-4,
                scorer_type="transe" if projection else "distmult", gat_out_dim=F_OUT, gat_heads=HEADS,
                gat_num_layers=LAYERS, dropout=0.0, rel_attn_dropout=0.0, project_to_input_size=projection,
                projection_layers=2, out_dir=str(tmp_path / ("b200" if patch else "ref")), num_neg=NUM_NEG,
                log_every_n_steps=10 ** 9, use_self_adv_neg=self_adv, self_adv_alpha=0.5, eval_ks_ranks=list(KS))
            losses = []
            calc = tr._calculate_loss

            def recording(**kw):
                out = calc(**kw)
                if kw.get("phase") == "train":
                    losses.append(float(out[2].detach()))
                return out

            tr._calculate_loss = recording
            tr.training_scheduler.prepare(epochs=1, train_dataset=tr.dataset.train_dataset,
                                          train_batch_size=tr.dataset.train_batch_size, optimizer=tr.optimizer)
            tr.single_epoch(epoch=1, epochs=1, epoch_loss=0.0, running_loss=0.0, running_examples=0)
            mrr, hits, eval_loss, *_ = tr.evaluate(ks=KS)
        kind = type(tr.model).__module__
        return dict(losses=losses, mrr=mrr, hits=hits, eval_loss=eval_loss, model_module=kind,
                    py_random_probe=random.random(), torch_probe=float(torch.rand(())))
    finally:
        for (m, n), v in saved.items():
            setattr(m, n, v)
        ref_trainer_mod.RelGATTrainer._calculate_loss = saved_calc


@pytest.mark.parametrize("projection,self_adv,fused_seam,receptive_field", [
    (False, False, False, None),   # DistMult, margin ranking loss: model.forward seam
    (True, True, False, None),     # the reference's shipped recipe: TransE + projection + self-adversarial + reconstruction
    (True, True, True, None),       # same, plus the trainer-level seam (fused batch rows, score and loss kernels)
    (False, True, True, None),
    (True, True, True, "masked"),   # ... and every train / eval batch restricted to its receptive field
    (False, False, False, "masked"),
    (True, True, True, "blocks"),   # ... through per-batch bipartite sub-indexes (blocks.py)
])
def test_reference_trainer_with_dropin_classes_matches_reference_on_cpu(dev, tmp_path, monkeypatch, projection, self_adv,
                                                                        fused_seam, receptive_field):
    ref = _run_trainer("cpu", tmp_path, projection=projection, self_adv=self_adv, patch=False)
    if receptive_field:
        monkeypatch.setenv("RELGAT_RECEPTIVE_FIELD", "1")  # read by the drop-in RelGATModel's constructor
        monkeypatch.setenv("RELGAT_RF_MODE", receptive_field)
    got = _run_trainer(dev, tmp_path, projection=projection, self_adv=self_adv, patch=True, fused_seam=fused_seam)
    assert ref["model_module"].startswith("relgat_projector.") and got["model_module"].startswith("relgat_projector_b200")
    assert len(ref["losses"]) == len(got["losses"]) >= 10
    err = np.abs(np.array(got["losses"]) - np.array(ref["losses"])) / np.maximum(1.0, np.abs(ref["losses"]))
    assert err.max() < LOSS_TOL, (err.max(), ref["losses"][:3], got["losses"][:3])
    assert abs(got["eval_loss"] - ref["eval_loss"]) < LOSS_TOL * max(1.0, abs(ref["eval_loss"]))
    assert abs(got["mrr"] - ref["mrr"]) < METRIC_TOL
    for k in KS:  # includes Hits@10 (num_neg = 10)
        assert abs(got["hits"][k] - ref["hits"][k]) < METRIC_TOL, k
    # both runs consumed the Python and torch RNG streams identically (same batches, same negatives)
    assert got["py_random_probe"] == ref["py_random_probe"] and got["torch_probe"] == ref["torch_probe"]
