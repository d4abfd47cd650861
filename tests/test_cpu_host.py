"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the drop-in modules reproduce the reference's parameter names / init RNG order, the
stream-exact sampler, the torch-side losses, and loud failure without a GPU."""
import os
import random

import numpy as np
import pytest
import torch

import relgat_projector_b200 as R
from oracle import relgat_oracle as O
from relgat_projector_b200 import _lib, loss as L
from relgat_projector_b200.batching import ReferenceStreamSampler, shuffle_and_split
from tests.helpers import GOLDEN_DIR, Case, MODEL_CASES


def test_library_loads_and_exports_header_symbols():
    lib = _lib.load()
    names = _lib.header_symbols()
    assert len(names) >= 16
    assert set(names) == set(_lib.SIGNATURES.keys())
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/relgat_b200.h but not exported"
    assert lib.relgat_abi_version() == _lib.ABI_VERSION == 13
    # host-side argument validation needs no GPU
    assert lib.relgat_graph_index_workspace_bytes(-1) == -1
    assert lib.relgat_gemm_workspace_bytes(128, 128, 64, 0, 0, 1) == 0
    assert lib.relgat_gemm_workspace_bytes(128, 64, 640, 1, 1, 4) == 4 * 128 * 64 * 4
    assert lib.relgat_layer_fwd(None, 0, 0, None, None, None, None, None, None, 0, None, 0, None, None, 0, None, None,
                                None, None, None, None, 0, None, None, None, None, None, 0, 1.0, None, 1.0, None,
                                1, 4, 1, 148, None, None) == -1
    assert lib.relgat_rank_loss(None, None, 1, 1, 1, 1, 0, 1.0, 1.0, 0, None, None, None, None) == -1
    assert lib.relgat_recon_loss(None, None, None, 1, 0, 4, 0, 0, 1.0, 1.0, 0.0, None, None, None, None, None, None) == -1
    assert lib.relgat_bernoulli_bits(None, 4, 0.5, 1, None) == -1
    assert lib.relgat_zero_rows(None, 4, None, 1, 4, None) == -1
    assert lib.relgat_score_fwd(7, 0, None, None, None, None, None, None, 1, 4, None, None, 0, None, None, None) == -1


def test_gemm_plan_host_logic(monkeypatch):
    """relgat_gemm_plan is pure host logic (no GPU): tile shapes the kernel can run, work units that cover the output,
    the measured choices of the config-2 shapes, and the experiment knobs."""
    from relgat_projector_b200 import ops
    for M, N, b_mn in [(300_000, 800, False), (1000, 1024, True), (1000, 800, True), (800, 1000, True), (200, 800, False),
                       (128, 64, False), (129, 16, False), (5120, 1024, False), (1_000_000, 2048, False), (513, 1000, True)]:
        cost, tm, tn, slots = ops.gemm_plan(M, N, b_mn, 148)
        assert tm == (256 if M > 128 else 128) and cost > 0
        assert tn % 16 == 0 and 16 <= tn <= 512 and slots == 148 // (tm // 128)
        if N <= 256:
            assert tn == (N + 15) // 16 * 16  # one N tile
    # 300k-row GEMMs: the N tile that moves the fewest operand bytes (4 x 208 rather than 5 x 160 or 4 x 256)
    assert ops.gemm_plan(300_000, 800, False, 148)[2] == 208
    # weight-gradient shapes (few row tiles): two 256-column tiles per work unit
    assert ops.gemm_plan(1000, 1024, True, 148)[2] == 512
    # both orientations of config 2's second-layer dW cost the same: no transposed product
    assert ops.gemm_cost_model(800, 1000) >= 0.97 * ops.gemm_cost_model(1000, 800)
    monkeypatch.setenv("RELGAT_GEMM_NTU", "1")
    assert ops.gemm_plan(1000, 1024, True, 148)[2] == 256
    monkeypatch.setenv("RELGAT_GEMM_CG", "1")
    assert ops.gemm_plan(1000, 1024, True, 148)[1:] == (128, 256, 148)
    monkeypatch.setenv("RELGAT_GEMM_BN", "160")
    assert ops.gemm_plan(300_000, 800, False, 148)[2] == 160
    assert _lib.load().relgat_gemm_plan(0, 5, 0, 148, None, None, None) == -1


def _build_model(c: Case, seed):
    torch.manual_seed(seed)
    return R.RelGATModel(
        node_emb=c.t("x0"), edge_index=c.edge_index(), edge_type=c.t("rel"), num_rel=c.r, scorer_type=c.scorer,
        gat_out_dim=c.f, gat_heads=c.h, dropout=0.0, relation_attn_dropout=0.0, gat_num_layers=c.layers,
        project_to_input_size=c.projection, projection_layers=c.proj_layers, projection_dropout=0.0,
        projection_hidden_dim=0).to(c.dtype)


@pytest.mark.parametrize("name,seed", [("tiny_fp64", 1), ("f200_fp32", 2), ("transe_proj_fp32", 3), ("transe_fp64", 5)])
def test_same_seed_same_parameters_and_state_dict_keys(name, seed):
    """Constructor consumes the RNG in the reference's order: identical weights,
    identical state-dict keys, strict load of a reference state dict."""
    c = Case(name)
    m = _build_model(c, seed)
    sd = m.state_dict()
    ref_keys = {k[len("param/"):] for k in c.z.files if k.startswith("param/")} | {"node_emb_fixed"}
    assert set(sd.keys()) == ref_keys
    for k in ref_keys - {"node_emb_fixed"}:
        if k.endswith("rel_bias"):
            continue  # the fixture overwrote the (zero-initialised) bias after construction
        assert np.array_equal(sd[k].numpy(), c.z["param/" + k]), k
    state = {k: torch.from_numpy(c.z["param/" + k]) for k in ref_keys - {"node_emb_fixed"}}
    state["node_emb_fixed"] = c.t("x0")
    m.load_state_dict(state, strict=True)


def test_single_layer_attribute_name():
    c = Case("f200_fp32")
    m = _build_model(c, 2)
    assert hasattr(m, "gat_layer") and not hasattr(m, "gat_layers") and m.act is None
    assert isinstance(m.edge_index, torch.Tensor) and "edge_index" not in dict(m.named_buffers())


def test_cpu_tensors_fail_loudly():
    c = Case("tiny_fp64")
    m = _build_model(c, 1).float()
    with pytest.raises(RuntimeError, match="CUDA device only"):
        m.single_gat_step()
    sc = R.DistMultScorer(3, 8)
    with pytest.raises(RuntimeError, match="CUDA device only"):
        sc(torch.randn(2, 8), torch.zeros(2, dtype=torch.long), torch.randn(2, 8))


def test_product_does_not_import_oracle():
    import relgat_projector_b200
    root = os.path.dirname(relgat_projector_b200.__file__)
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            text = open(os.path.join(root, fn), encoding="utf-8").read()
            assert "import oracle" not in text and "from oracle" not in text, fn


def test_stream_exact_sampler_matches_reference_fixture():
    z = np.load(os.path.join(GOLDEN_DIR, "sampling.npz"))
    n, t, r, d, k, bs, seed = [int(v) for v in z["meta"]]
    raw = list(zip(z["raw_src"].tolist(), z["raw_dst"].tolist(), z["raw_rel"].tolist()))
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    train, ev = shuffle_and_split(raw, 0.9)
    assert np.array_equal(np.array([e[0] for e in train]), z["edge_index"][0])
    assert np.array_equal(np.array(ev), z["eval_edges"])
    for bi, (s, rr, dd) in enumerate(ReferenceStreamSampler(train, n, k, bs)):
        if bi >= 3:
            break
        assert np.array_equal(s.numpy(), z[f"batch{bi}_src"])
        assert np.array_equal(rr.numpy(), z[f"batch{bi}_rel"])
        assert np.array_equal(dd.numpy(), z[f"batch{bi}_dst"])
        assert s.dtype == torch.int64 and s.numel() == bs * (1 + k)


def test_losses_and_metric_match_oracle():
    g = torch.Generator().manual_seed(0)
    b, k, d = 17, 5, 12
    pos = torch.randn(b, generator=g, dtype=torch.float64)
    neg = torch.randn(b, k, generator=g, dtype=torch.float64)
    tr, dv = torch.randn(b, d, generator=g, dtype=torch.float64), torch.randn(b, d, generator=g, dtype=torch.float64)
    ndv = torch.randn(k, b, d, generator=g, dtype=torch.float64)
    margin = L.RelGATLoss("margin", None, 0.7, None, {})
    sadv = L.RelGATLoss("self_adversarial_loss", 0.9, None, None, {})
    assert torch.equal(margin.prepare_scores_and_compute_loss(pos, neg), O.margin_ranking_loss_port(pos, neg, 0.7))
    assert torch.equal(sadv.prepare_scores_and_compute_loss(pos, neg), O.self_adversarial_loss_port(pos, neg, 0.9))
    for w in [(1.0, 1.0, 1.0, 0.0), (1.0, 0.0, 2.0, 0.5), (0.0, 1.0, 0.0, 0.0)]:
        mo = L.MultiObjectiveRelLoss(relgat_loss=margin, run_config={}, relgat_weight=w[0], pos_cosine_weight=w[1],
                                     neg_cosine_weight=w[2], mse_weight=w[3])
        a = mo(pos_score=pos, neg_score=neg, transformed_src=tr, dst_vec=dv, neg_dst_vec=ndv)
        e = O.multi_objective_loss_port(pos, neg, tr, dv, ndv, ranking_loss=lambda p, n: O.margin_ranking_loss_port(p, n, 0.7),
                                        w_rank=w[0], w_pos=w[1], w_neg=w[2], w_mse=w[3])
        assert torch.allclose(a, e, rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        L.MultiObjectiveRelLoss(relgat_loss=margin, run_config={}, relgat_weight=0.0, pos_cosine_weight=0.0,
                                neg_cosine_weight=0.0, mse_weight=0.0)(pos_score=pos, neg_score=neg, transformed_src=tr,
                                                                       dst_vec=dv, neg_dst_vec=ndv)
    neg2 = neg.clone(); neg2[0, 0] = float("nan"); neg2[1, 1] = float("inf"); neg2[2, 2] = pos[2]
    m1, h1 = L.compute_mrr_hits(pos, neg2, ks=(1, 3, 5))
    m2, h2 = O.mrr_hits_port(pos, neg2, (1, 3, 5))
    assert m1 == m2 and h1 == h2
    assert L.compute_mrr_hits(pos[:0], neg[:0], ks=(1,)) == (0.0, {1: 0.0})
    flat = torch.arange(b * (1 + k), dtype=torch.float64)
    p1, n1 = L.split_scores(flat, b, k)
    p2, n2 = O.split_scores_kmajor(flat, b, k)
    assert torch.equal(n1, n2) and torch.equal(p1, p2)
    assert torch.equal(L.split_scores(flat, b, k, projection_path=True)[1], O.split_scores_projection_path(flat, b, k)[1])


def test_synthetic_graph_is_seeded_and_well_formed():
    from relgat_projector_b200 import synthetic as S
    a = S.tensor_kg(500, 3000, 7, 8, seed=3)
    b = S.tensor_kg(500, 3000, 7, 8, seed=3)
    assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.node_emb, b.node_emb)
    assert a.edge_index.shape == (2, 2700) and a.eval_triples.shape == (300, 3)
    assert int((a.edge_index[0] == a.edge_index[1]).sum()) == 0
    assert int(a.edge_type.max()) < 7
    sk = S.tensor_kg(500, 3000, 7, 8, seed=3, skew=1.1)
    deg = torch.bincount(sk.edge_index[1], minlength=500)
    assert int(deg.max()) > 10 * int(torch.bincount(a.edge_index[1], minlength=500).max()) // 4
    g = torch.Generator().manual_seed(1)
    s, r, d = S.sample_batch(a.train_triples, 500, 16, 3, g)
    assert s.shape == (64,) and torch.equal(s[:16], s[16:32]) and torch.equal(r[:16], r[48:])
    assert int((d[16:].view(3, 16) == d[:16].unsqueeze(0)).sum()) == 0
    n2e, r2i, raw = S.reference_inputs(50, 200, 4, 6, seed=1)
    assert len(n2e) == 50 and len(raw) == 200 and raw[0][2].startswith("rel_")


def test_native_sampler_is_bit_exact_with_cpython_stream():
    """C restatement of random.choice / random.shuffle: same ids, same generator state afterwards,
    for node counts around powers of two (rejection paths) and against the reference fixture."""
    from relgat_projector_b200.batching import NativeStreamSampler, native_shuffle_and_split
    for n_nodes in (2, 3, 200, 255, 256, 257, 65536, 300_000, 5_000_000):
        edges = [(i % n_nodes, (7 * i + 1) % n_nodes, i % 5) for i in range(300)]
        random.seed(99); torch.manual_seed(99)
        py = ReferenceStreamSampler(edges, n_nodes, 6, 64, shuffle=False)
        want = [py.build(list(range(b * 64, b * 64 + 64))) for b in range(4)]
        tail_py = random.random()
        random.seed(99)
        nat = NativeStreamSampler(edges, n_nodes, 6, 64, shuffle=False)
        got = [nat.build(list(range(b * 64, b * 64 + 64))) for b in range(4)]
        assert random.random() == tail_py  # generator left in the identical state
        for w, g in zip(want, got):
            for a, b_ in zip(w, g):
                assert torch.equal(a, b_)
    z = np.load(os.path.join(GOLDEN_DIR, "sampling.npz"))
    n, t, r, d, k, bs, seed = [int(v) for v in z["meta"]]
    raw = list(zip(z["raw_src"].tolist(), z["raw_dst"].tolist(), z["raw_rel"].tolist()))
    random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
    train, ev = native_shuffle_and_split(raw, 0.9)
    assert np.array_equal(np.array([e[0] for e in train]), z["edge_index"][0])
    assert np.array_equal(np.array(ev), z["eval_edges"])
    for bi, (s_, rr, dd) in enumerate(NativeStreamSampler(train, n, k, bs)):
        if bi >= 3:
            break
        assert np.array_equal(s_.numpy(), z[f"batch{bi}_src"])
        assert np.array_equal(dd.numpy(), z[f"batch{bi}_dst"])


def test_sparse_row_tag_is_dropped_after_in_place_edit():
    """functional.mark_sparse_rows / sparse_rows_of: the row list travels with the gradient tensor and is ignored
    as soon as anything (e.g. autograd's in-place accumulation of a second gradient) has written to the tensor."""
    from relgat_projector_b200 import functional as Fn
    g = torch.zeros(5, 3)
    assert Fn.sparse_rows_of(g) is None
    Fn.mark_sparse_rows(g, torch.tensor([1, 3]))
    assert Fn.sparse_rows_of(g).tolist() == [1, 3]
    g.add_(1.0)
    assert Fn.sparse_rows_of(g) is None


def test_save_and_load_pretrained_round_trip(tmp_path):
    """(f)-4: the reference's get_config reads an attribute it never sets (model.py:188-194), so its own
    save_pretrained cannot run; here config.json + pytorch_model.bin round-trip, with the reference's file names and
    state-dict keys (a checkpoint written by the reference trainer's torch.save(state_dict) loads the same way)."""
    from relgat_projector_b200 import synthetic as S
    kg = S.tensor_kg(60, 300, 4, 12, seed=3)
    torch.manual_seed(5)
    m = R.RelGATModel(kg.node_emb, kg.edge_index, kg.edge_type, num_rel=4, scorer_type="transe", gat_out_dim=6,
                      gat_heads=2, dropout=0.1, gat_num_layers=2, project_to_input_size=True, projection_layers=2)
    out = tmp_path / "ckpt"
    m.save_pretrained(str(out), add_files=[("run.json", {"note": "ok"})])
    assert sorted(os.listdir(out)) == ["config.json", "pytorch_model.bin", "run.json"]
    m2 = R.RelGATModel.load_from_pretrained(str(out), node_emb=kg.node_emb, edge_index=kg.edge_index,
                                            edge_type=kg.edge_type)
    assert m2.get_config() == m.get_config() and not m2.training
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd) == list(sd2) and all(torch.equal(sd[k], sd2[k]) for k in sd)
    with pytest.raises(ValueError):
        R.RelGATModel.load_from_pretrained(str(out), node_emb=kg.node_emb[:, :5], edge_index=kg.edge_index,
                                           edge_type=kg.edge_type)
    with pytest.raises(FileNotFoundError):
        R.RelGATModel.load_from_pretrained(str(tmp_path / "missing"), node_emb=kg.node_emb)


def test_node_table_matches_reference_matrix_and_is_memory_mapped(tmp_path):
    """storage.py: the memory-mapped node table equals the matrix the reference builds from the pickled dict
    (torch.stack over sorted ids, dataset/relgat_dataset.py:61-68), bit for bit."""
    import pickle
    from relgat_projector_b200 import storage
    rng = np.random.default_rng(0)
    node2emb = {}
    for nid in rng.permutation(50)[:37] * 3 + 11:  # unordered, sparse ids; mixed containers like the reference accepts
        v = rng.standard_normal(9).astype(np.float32)
        node2emb[int(nid)] = [v.tolist(), v, torch.from_numpy(v.copy())][int(nid) % 3]
    ref_ids = sorted(node2emb.keys())
    ref = torch.stack([torch.as_tensor(node2emb[nid]) for nid in ref_ids], dim=0).to(torch.float32)
    pk = tmp_path / "nodes.pkl"
    with open(pk, "wb") as f:
        pickle.dump({k: (v.numpy() if torch.is_tensor(v) else v) for k, v in node2emb.items()}, f)
    assert storage.convert_pickled_nodes(str(pk), str(tmp_path / "tab")) == (37, 9)
    ids, emb = storage.load_node_table(str(tmp_path / "tab"))
    assert ids.tolist() == ref_ids and torch.equal(emb, ref)
    assert storage.id_to_row(ids) == {nid: i for i, nid in enumerate(ref_ids)}
    emb[0, 0] += 1.0  # copy-on-write: the file stays as written
    _, again = storage.load_node_table(str(tmp_path / "tab"), mmap=False)
    assert torch.equal(again, ref)
    with pytest.raises(ValueError):
        storage.write_node_table({1: [1.0, 2.0], 2: [1.0]}, str(tmp_path / "bad"))


def test_checkpoint_without_the_frozen_node_table(tmp_path):
    """storage.save_trainable_state leaves the [N, D_in] input buffer out of the checkpoint (the reference stores it
    in every one) and load_trainable_state restores every trainable tensor, checking the table's fingerprint."""
    from relgat_projector_b200 import storage
    torch.manual_seed(0)
    x = torch.randn(300, 64)
    ei = torch.randint(0, 300, (2, 900))
    et = torch.randint(0, 5, (900,))
    kw = dict(num_rel=5, gat_out_dim=8, gat_heads=2, gat_num_layers=2, project_to_input_size=True, projection_layers=2)
    m = R.RelGATModel(x, ei, et, **kw)
    full = tmp_path / "full.pt"
    torch.save(m.state_dict(), full)
    small = storage.save_trainable_state(m, str(tmp_path / "small.pt"))
    assert small < os.path.getsize(full) - x.numel() * 4 + 4096
    torch.manual_seed(1)
    m2 = R.RelGATModel(x.clone(), ei, et, **kw)
    storage.load_trainable_state(m2, str(tmp_path / "small.pt"))
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    storage.load_trainable_state(m2, str(full))  # the reference's own format loads too
    m3 = R.RelGATModel(x + 1.0, ei, et, **kw)
    with pytest.raises(ValueError):
        storage.load_trainable_state(m3, str(tmp_path / "small.pt"))
    m4 = R.RelGATModel(x, ei, et, **dict(kw, gat_heads=1))
    with pytest.raises((KeyError, RuntimeError)):
        storage.load_trainable_state(m4, str(tmp_path / "small.pt"))
