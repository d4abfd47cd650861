"""GPU parity tests of the round-2 additions: fused ranking / reconstruction loss kernels, dropout inside the edge
kernels (injected masks replayed in the oracle), the stack + batch-row gather node (no dense [N, C] gradient), and
parity at the sizes of BASELINE.json's configs (c1 in full against the oracle port; c2 through sampled destination
rows and the parameter gradients they induce)."""
import numpy as np
import pytest
import torch

import relgat_projector_b200 as R
from oracle import relgat_oracle as O
from relgat_projector_b200 import functional as RF, loss as L, ops, synthetic as S
from relgat_projector_b200.graph import GraphIndex
from tests.helpers import Case, MODEL_CASES, rel_err

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4  # north_star tolerance (fp32): logits / embeddings / losses / gradients


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _load_model(c: Case, dev):
    m = R.RelGATModel(
        node_emb=c.t("x0").float().to(dev), edge_index=c.edge_index().to(dev), edge_type=c.t("rel").to(dev),
        num_rel=c.r, scorer_type=c.scorer, gat_out_dim=c.f, gat_heads=c.h, dropout=0.0, relation_attn_dropout=0.0,
        gat_num_layers=c.layers, project_to_input_size=c.projection, projection_layers=c.proj_layers,
        projection_dropout=0.0, projection_hidden_dim=0).to(dev)
    state = {k[len("param/"):]: torch.from_numpy(c.z[k]).float() for k in c.z.files if k.startswith("param/")}
    state["node_emb_fixed"] = c.t("x0").float()
    m.load_state_dict(state, strict=True)
    return m.train()


# ------------------------------------------------------------------------------------------------------
# the trainer-level fused path against the reference's golden outputs
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", MODEL_CASES)
def test_fused_calculate_loss_matches_reference_golden(dev, name):
    """loss.calculate_loss = stack + row gather -> [projection of the batch rows] -> one score launch -> fused loss
    kernels, checked against the values and parameter gradients the UNMODIFIED reference produced (tests/golden)."""
    c = Case(name)
    m = _load_model(c, dev)
    rank = L.RelGATLoss("self_adversarial_loss" if c.loss_type == "self_adv" else "margin", 0.7, 1.0, None, {})
    multi = L.MultiObjectiveRelLoss(relgat_loss=rank, run_config={}, relgat_weight=c.weights[0],
                                    pos_cosine_weight=c.weights[1], neg_cosine_weight=c.weights[2],
                                    mse_weight=c.weights[3]) if c.projection else None
    ids = [c.t(k).to(dev) for k in ("src_ids", "rel_ids", "dst_ids")]
    pos, neg, loss, mse, cpos, cneg = L.calculate_loss(m, *ids, c.b, rank, multi)
    assert rel_err(pos.detach().cpu().numpy(), c.z["pos"]) < FP32_TOL
    assert rel_err(neg.detach().cpu().numpy(), c.z["neg"]) < FP32_TOL
    assert abs(float(loss.detach()) - float(c.z["loss"])) < FP32_TOL * max(1.0, abs(float(c.z["loss"])))
    loss.backward()
    for pname, p in m.named_parameters():
        ref = c.z["grad/" + pname]
        assert p.grad is not None, pname
        tol = 5 * FP32_TOL if ref.size and np.abs(ref).max() < 1e-6 else FP32_TOL
        assert rel_err(p.grad.cpu().numpy(), ref) < tol, pname
    if c.projection:
        assert all(torch.isfinite(v) for v in (mse, cpos, cneg))
    # the zero table the backward borrowed is all-zero again
    for pool in RF._ZERO_TABLES.values():
        for t in pool:
            assert not bool(t.any())


# ------------------------------------------------------------------------------------------------------
# loss kernels
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["margin", "self_adversarial_loss"])
@pytest.mark.parametrize("layout", ["kmajor", "projection"])
def test_rank_loss_kernel_vs_oracle(dev, kind, layout):
    g = torch.Generator().manual_seed(3)
    b, k = 37, 5
    flat = (torch.randn(b * (1 + k), generator=g) * 3).float()
    flat[5], flat[b + 7], flat[b + 11] = float("nan"), float("inf"), float("-inf")
    ref_in = flat.double().clone().requires_grad_(True)
    clean = torch.nan_to_num(ref_in, nan=0.0, neginf=-1e9, posinf=1e9)
    rp, rn = (O.split_scores_kmajor if layout == "kmajor" else O.split_scores_projection_path)(clean, b, k)
    ref = O.margin_ranking_loss_port(rp, rn, 0.8) if kind == "margin" else O.self_adversarial_loss_port(rp, rn, 0.6)
    ref.backward()
    x = flat.to(dev).requires_grad_(True)
    pos = x[:b]
    neg = x[b:].view(k, b).t() if layout == "kmajor" else x[b:].view(b, k)  # strided views, no copies
    rl = L.RelGATLoss(kind, 0.6, 0.8, None, {})
    loss = rl.prepare_scores_and_compute_loss(pos, neg, sanitize=True)
    loss.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert rel_err(x.grad.cpu().numpy(), ref_in.grad.numpy()) < 1e-5
    # bitwise reproducible
    loss2 = rl.prepare_scores_and_compute_loss(pos.detach(), neg.detach(), sanitize=True)
    assert torch.equal(loss.detach(), loss2)
    # finite scores: same value without the sanitising flag
    fin = torch.nan_to_num(flat, nan=0.0, neginf=-5.0, posinf=5.0).to(dev)
    a = rl.prepare_scores_and_compute_loss(fin[:b], fin[b:].view(b, k))
    bb = rl.prepare_scores_and_compute_loss(fin[:b], fin[b:].view(b, k), sanitize=True)
    assert torch.equal(a, bb)


@pytest.mark.parametrize("weights", [(1.0, 1.0, 0.0), (0.5, 2.0, 0.3), (0.0, 1.0, 1.0)])
def test_recon_loss_kernel_vs_oracle(dev, weights):
    g = torch.Generator().manual_seed(5)
    b, k, d = 19, 4, 72
    tr = torch.randn(b, d, generator=g)
    dst = torch.randn(b, d, generator=g)
    flat_nd = torch.randn(b * k, d, generator=g)  # the trainer's flat negative rows; (b, k) = row b*k + k
    tr[3] = 0.0  # F.normalize eps branch
    rt, rd, rn = (t.double().clone().requires_grad_(True) for t in (tr, dst, flat_nd))
    ndv = rn.view(b, k, d).permute(1, 0, 2).contiguous()
    w_pos, w_neg, w_mse = weights
    parts = []
    if w_pos:
        parts.append(w_pos * O.cosine_loss_port(rt, rd))
    if w_neg:
        parts.append(w_neg * (1.0 - O.cosine_loss_port(rt, ndv)))
    if w_mse:
        parts.append(w_mse * torch.nn.functional.mse_loss(rt, rd))
    ref = torch.stack(parts).sum()
    ref.backward()
    xt, xd, xn = (t.to(dev).requires_grad_(True) for t in (tr, dst, flat_nd))
    view = xn.view(b, k, d).permute(1, 0, 2)  # [K, B, D] strided view: no copy
    total, values = RF.fused_recon_loss(xt, xd, view, w_pos, w_neg, w_mse)
    total.backward()
    assert abs(float(total) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    with torch.no_grad():
        want = [float(O.cosine_loss_port(rt, rd)), float(O.cosine_loss_port(rt, ndv)),
                float(torch.nn.functional.mse_loss(rt, rd))]
    assert np.allclose(values.cpu().numpy(), want, rtol=1e-5, atol=1e-6)
    for got, exp in ((xt.grad, rt.grad), (xd.grad, rd.grad), (xn.grad, rn.grad)):
        if exp is None:
            assert got is None or not bool(got.any())
            continue
        assert rel_err(got.cpu().numpy(), exp.numpy()) < 1e-5
    # MultiObjectiveRelLoss class, CUDA inputs -> the same kernels
    rank = L.RelGATLoss("margin", None, 1.0, None, {})
    multi = L.MultiObjectiveRelLoss(relgat_loss=rank, run_config={}, relgat_weight=1.0, pos_cosine_weight=w_pos,
                                    neg_cosine_weight=w_neg, mse_weight=w_mse)
    pos = torch.randn(b, generator=g)
    neg = torch.randn(b, k, generator=g)
    got = multi(pos_score=pos.to(dev), neg_score=neg.to(dev), transformed_src=tr.to(dev), dst_vec=dst.to(dev),
                neg_dst_vec=flat_nd.view(b, k, d).permute(1, 0, 2).contiguous().to(dev))
    exp = O.multi_objective_loss_port(pos.double(), neg.double(), tr.double(), dst.double(), ndv.detach(),
                                      ranking_loss=lambda p, n: O.margin_ranking_loss_port(p, n, 1.0),
                                      w_rank=1.0, w_pos=w_pos, w_neg=w_neg, w_mse=w_mse)
    assert abs(float(got) - float(exp)) < 1e-5 * max(1.0, abs(float(exp)))


def test_scorer_shape_errors_raise(dev):
    """A wrong-width operand must raise like torch would in the reference (scorer.py:80-83), not read out of bounds."""
    sc = R.DistMultScorer(5, 16).to(dev)
    rel = torch.zeros(4, dtype=torch.long, device=dev)
    with pytest.raises(ValueError):
        sc(torch.randn(4, 24, device=dev), rel, torch.randn(4, 24, device=dev))
    with pytest.raises(ValueError):
        sc(torch.randn(3, 16, device=dev), rel, torch.randn(4, 16, device=dev))
    with pytest.raises(ValueError):
        sc.gather_score(torch.randn(9, 16, device=dev), rel[:3], rel, rel)
    assert sc(torch.randn(4, 16, device=dev), rel, torch.randn(4, 16, device=dev)).shape == (4,)


# ------------------------------------------------------------------------------------------------------
# dropout inside the edge kernels
# ------------------------------------------------------------------------------------------------------
def test_bernoulli_bits_rate_and_seeding(dev):
    p = 0.3
    m = ops.DropMask.draw((4096, 25), p, dev, seed=1234)
    ones = sum(int(((m.bits >> s) & 1).sum()) for s in range(32))
    rate = ones / (4096 * 25 * 32)
    assert abs(rate - (1 - p)) < 2e-3 and m.scale == pytest.approx(1 / (1 - p))
    assert torch.equal(m.bits, ops.DropMask.draw((4096, 25), p, dev, seed=1234).bits)
    assert not torch.equal(m.bits, ops.DropMask.draw((4096, 25), p, dev, seed=1235).bits)
    torch.manual_seed(7)
    a = ops.DropMask.draw((64, 3), p, dev).bits
    torch.manual_seed(7)
    assert torch.equal(a, ops.DropMask.draw((64, 3), p, dev).bits)  # torch.manual_seed governs the masks
    keep = torch.rand(50, 70, device=dev) > 0.5
    packed = ops.DropMask.feature_mask(keep, 0.5).bits
    cols = torch.arange(70, device=dev)
    back = ((packed[:, cols // 32] >> (cols % 32)) & 1).bool()
    assert torch.equal(back, keep)


@pytest.mark.parametrize("sites", ["feature", "attention", "both"])
def test_dropout_inside_kernels_matches_oracle_with_injected_masks(dev, sites):
    """Training-mode dropout (reference layer.py:296-297, 321-322; p = 0.3 / 0.2 in its scripts) runs inside the fused
    stack.  The masks are injected, the oracle port replays them: outputs and every gradient agree at 1e-4."""
    rng = np.random.default_rng(11)
    n, e, r, d_in, f, h, L_ = 300, 2400, 6, 48, 24, 4, 2
    p_feat, p_attn = 0.3, 0.2
    x0 = torch.from_numpy(rng.standard_normal((n, d_in)).astype(np.float32))
    ei = torch.from_numpy(np.stack([rng.integers(0, n, e), rng.integers(0, n - 10, e)]).astype(np.int64))
    et = torch.from_numpy(rng.integers(0, r, e).astype(np.int64))
    torch.manual_seed(1)
    layers = [R.RelGATLayer(d_in if l == 0 else h * f, f, r, heads=h, dropout=0.0).to(dev) for l in range(L_)]
    for lyr in layers:
        torch.nn.init.normal_(lyr.rel_bias, std=0.1)
    g = GraphIndex(ei.to(dev), et.to(dev), n, r)
    perm = g.csr_perm.long().cpu()
    drops, omasks = [], []
    for l in range(L_):
        fk = torch.from_numpy(rng.random((n, h * f)) >= p_feat) if sites in ("feature", "both") else None
        ak = torch.from_numpy(rng.random((e, h)) >= p_attn) if sites in ("attention", "both") else None  # COO order
        drops.append(RF.LayerDropout(ops.DropMask.feature_mask(fk.to(dev), p_feat) if fk is not None else None,
                                     ops.DropMask.edge_mask(ak[perm].to(dev), p_attn) if ak is not None else None))
        omasks.append(dict(feat_keep=fk, feat_p=p_feat if fk is not None else 0.0,
                           attn_keep=ak, attn_p=p_attn if ak is not None else 0.0))
    ids = torch.from_numpy(rng.integers(0, n, 80))
    gout = torch.from_numpy(rng.standard_normal((80, h * f)).astype(np.float32))
    rows = RF.relgat_stack(x0.to(dev), g, h, f, [lyr.kernel_params() for lyr in layers], drop=drops,
                           gather_ids=ids.to(dev))
    rows.backward(gout.to(dev))
    ol = [{"W": [p.weight.detach().cpu().double().requires_grad_(True) for p in lyr.proj],
           "A": [a.detach().cpu().double().requires_grad_(True) for a in lyr.attn_vec],
           "beta": lyr.rel_bias.detach().cpu().double().requires_grad_(True)} for lyr in layers]
    ref = O.gat_stack_port(x0.double(), ol, ei, et, masks=omasks)[ids]
    ref.backward(gout.double())
    assert rel_err(rows.detach().cpu().numpy(), ref.detach().numpy()) < FP32_TOL
    for lyr, o in zip(layers, ol):
        for hh in range(h):
            assert rel_err(lyr.proj[hh].weight.grad.cpu().numpy(), o["W"][hh].grad.numpy()) < FP32_TOL
            assert rel_err(lyr.attn_vec[hh].grad.cpu().numpy(), o["A"][hh].grad.numpy()) < FP32_TOL
        assert rel_err(lyr.rel_bias.grad.cpu().numpy(), o["beta"].grad.numpy()) < FP32_TOL


def test_training_with_reference_default_dropout_uses_one_fused_node(dev):
    c = Case("tiny_fp64")
    m = R.RelGATModel(node_emb=c.t("x0").float().to(dev), edge_index=c.edge_index().to(dev),
                      edge_type=c.t("rel").to(dev), num_rel=c.r, gat_out_dim=c.f, gat_heads=c.h, dropout=0.3,
                      relation_attn_dropout=0.1, gat_num_layers=2).to(dev).train()
    x = m.single_gat_step()
    assert type(x.grad_fn).__name__ == "RelGATStackFunctionBackward"  # no per-layer nodes, no ATen dropout
    frac0 = float((x == 0).float().mean())
    assert 0.2 < frac0 < 0.45
    torch.manual_seed(3)
    a = m.single_gat_step()
    torch.manual_seed(3)
    assert torch.equal(a, m.single_gat_step())  # masks follow torch's seed
    x.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.gat_layers.parameters())


# ------------------------------------------------------------------------------------------------------
# stack + batch-row gather
# ------------------------------------------------------------------------------------------------------
def test_stack_gather_equals_dense_stack_then_index(dev):
    c = Case("f200_fp32")
    ids = torch.cat([c.t("src_ids"), c.t("dst_ids")]).to(dev)
    gout = torch.randn(ids.numel(), c.h * c.f, generator=torch.Generator().manual_seed(0)).to(dev)
    res = []
    for fused in (True, False):
        m = _load_model(c, dev)
        rows = m._stack_output(gather_ids=ids) if fused else m._stack_output()[ids]
        rows.backward(gout)
        res.append((rows.detach(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}))
    assert torch.equal(res[0][0], res[1][0])
    assert res[0][1].keys() == res[1][1].keys()
    for nme in res[0][1]:  # torch's index backward sums duplicate rows with atomics: order differs, values agree
        assert rel_err(res[0][1][nme].cpu().numpy(), res[1][1][nme].cpu().numpy()) < 1e-5, nme
    # a second backward through the released state fails loudly instead of with a TypeError on None
    m = _load_model(c, dev)
    rows = m._stack_output(gather_ids=ids)
    rows.sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="released"):
        rows.sum().backward()


# ------------------------------------------------------------------------------------------------------
# parity at the sizes of the named configs
# ------------------------------------------------------------------------------------------------------
def _oracle_layers(model):
    return [{"W": [p.weight.detach().cpu().double().requires_grad_(True) for p in lyr.proj],
             "A": [a.detach().cpu().double().requires_grad_(True) for a in lyr.attn_vec],
             "beta": lyr.rel_bias.detach().cpu().double().requires_grad_(True)} for lyr in model._layers()]


def test_config1_full_size_training_step_vs_oracle(dev):
    """BASELINE.json configs[0] in full (10 k nodes / 50 k triplets -> 45 k message-passing edges, 1024-d, 2 layers,
    4 heads x 200, DistMult, B = 256, K = 4): loss, sampled rows of the final embeddings and EVERY parameter gradient
    against the oracle port of the reference's op sequence (fp64 on the CPU), 1e-4 relative."""
    cfg = S.CONFIGS["c1"]
    kg = S.tensor_kg(cfg["N"], cfg["T"], cfg["R"], cfg["D_in"], seed=42, device="cpu")
    torch.manual_seed(42)
    m = R.RelGATModel(kg.node_emb.to(dev), kg.edge_index.to(dev), kg.edge_type.to(dev), num_rel=cfg["R"],
                      scorer_type="distmult", gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=0.0,
                      gat_num_layers=cfg["L"]).to(dev).train()
    with torch.no_grad():
        for lyr in m._layers():
            torch.nn.init.normal_(lyr.rel_bias, std=0.05)  # the reference initialises it to zero: make the term count
    b, k = cfg["B"], cfg["K"]
    src, rel, dst = S.sample_batch(kg.train_triples, cfg["N"], b, k, torch.Generator().manual_seed(1))
    rank = L.RelGATLoss("margin", None, 1.0, None, {})
    pos, neg, loss, *_ = L.calculate_loss(m, src.to(dev), rel.to(dev), dst.to(dev), b, rank)
    loss.backward()
    with torch.no_grad():
        xf = m.eval().single_gat_step()
    layers = _oracle_layers(m)
    rel_emb = m.scorer.rel_emb.weight.detach().cpu().double().requires_grad_(True)
    ref_loss, rpos, rneg = O.train_step_port(kg.node_emb.double(), layers, rel_emb, kg.edge_index, kg.edge_type, src,
                                             rel, dst, scorer="distmult", b=b, k=k, margin=1.0)
    ref_loss.backward()
    with torch.no_grad():
        ref_x = O.gat_stack_port(kg.node_emb.double(), layers, kg.edge_index, kg.edge_type)
    assert abs(float(loss) - float(ref_loss)) < FP32_TOL * max(1.0, abs(float(ref_loss)))
    assert rel_err(pos.detach().cpu().numpy(), rpos.detach().numpy()) < FP32_TOL
    assert rel_err(neg.detach().cpu().numpy(), rneg.detach().numpy()) < FP32_TOL
    assert rel_err(xf.cpu().numpy(), ref_x.numpy()) < FP32_TOL
    iso = (torch.bincount(kg.edge_index[1], minlength=cfg["N"]) == 0).to(dev)
    assert bool(iso.any()) and not bool(xf[iso].any())  # nodes without in-edges: exactly zero rows
    for lyr, o in zip(m._layers(), layers):
        for hh in range(cfg["H"]):
            assert rel_err(lyr.proj[hh].weight.grad.cpu().numpy(), o["W"][hh].grad.numpy()) < FP32_TOL
            assert rel_err(lyr.attn_vec[hh].grad.cpu().numpy(), o["A"][hh].grad.numpy()) < FP32_TOL
        assert rel_err(lyr.rel_bias.grad.cpu().numpy(), o["beta"].grad.numpy()) < FP32_TOL
    assert rel_err(m.scorer.rel_emb.weight.grad.cpu().numpy(), rel_emb.grad.numpy()) < FP32_TOL


def test_config2_size_layer_sampled_rows_and_gradients(dev):
    """The 148-CTA persistent path at BASELINE.json configs[1] size (300 k nodes / 1.35 M edges, 1024-d, 4 heads x 200):
    1 500 sampled destination rows of the layer output, and the parameter gradients induced by a gradient that lives on
    those rows, against the fp64 closed form evaluated on the sampled rows' in-neighbourhood only."""
    cfg = S.CONFIGS["c2"]
    N, R_, H, F, D = cfg["N"], cfg["R"], cfg["H"], cfg["F"], cfg["D_in"]
    kg = S.tensor_kg(N, cfg["T"], R_, D, seed=42, device=str(dev))
    torch.manual_seed(0)
    layer = R.RelGATLayer(D, F, R_, heads=H, dropout=0.0).to(dev)
    with torch.no_grad():
        torch.nn.init.normal_(layer.rel_bias, std=0.05)
    out = layer(kg.node_emb, kg.edge_index, kg.edge_type)
    rng = np.random.default_rng(0)
    samp = np.sort(rng.choice(N, 1500, replace=False))
    gs = rng.standard_normal((1500, H * F)).astype(np.float32)
    dY = torch.zeros_like(out)
    dY[torch.from_numpy(samp).to(dev)] = torch.from_numpy(gs).to(dev)
    out.backward(dY)
    # oracle on the in-neighbourhood of the sampled rows (original edge order preserved inside a destination)
    src, dst = kg.edge_index[0].cpu().numpy(), kg.edge_index[1].cpu().numpy()
    rel = kg.edge_type.cpu().numpy()
    keep = np.isin(dst, samp)
    s_e, d_e, r_e = src[keep], dst[keep], rel[keep]
    nodes = np.unique(np.concatenate([s_e, samp]))
    loc = {int(v): i for i, v in enumerate(nodes)}
    ls = np.array([loc[int(v)] for v in s_e])
    ld = np.array([loc[int(v)] for v in d_e])
    X = kg.node_emb[torch.from_numpy(nodes).to(dev)].cpu().double().numpy()
    W = layer.packed_weight().detach().cpu().double().numpy()          # [H*F, D]
    A = layer.packed_attention().detach().cpu().double().numpy()       # [H, R, F]
    beta = layer.rel_bias.detach().cpu().double().numpy()
    P = (X @ W.T).reshape(len(nodes), H, F)
    gi = O.graph_index_np(ls, ld, r_e, len(nodes), R_)
    o_ref, z, alpha, _ = O.layer_forward_closed(P, A, beta, gi)
    sl = np.array([loc[int(v)] for v in samp])
    got = out.detach()[torch.from_numpy(samp).to(dev)].cpu().numpy()
    assert rel_err(got, o_ref[sl].reshape(len(samp), H * F)) < FP32_TOL
    G = np.zeros((len(nodes), H, F))
    G[sl] = gs.reshape(-1, H, F)
    dP, dA, dbeta, _ = O.layer_backward_closed(G, P, A, gi, z, alpha)
    dW = dP.reshape(len(nodes), H * F).T @ X
    assert rel_err(torch.stack([a.grad for a in layer.attn_vec]).cpu().numpy(), dA) < FP32_TOL
    assert rel_err(layer.rel_bias.grad.cpu().numpy(), dbeta) < FP32_TOL
    assert rel_err(torch.cat([p.weight.grad for p in layer.proj]).cpu().numpy(), dW) < FP32_TOL


@pytest.mark.parametrize("with_drop", [False, True])
def test_dx_gemm_with_fused_backward_prep(dev, with_drop):
    """relgat_gemm_dx_prep: the dX GEMM whose epilogue applies ELU'(y), the dropout mask and the per-head row sums of
    the layer below, against the unfused pair (GEMM, then relgat_layer_bwd_prep on its output)."""
    g = torch.Generator(device=dev).manual_seed(2)
    M, H, F, K = 1000, 4, 200, 800
    N = H * F
    assert ops.gemm_dx_prep_supported(N, F) and not ops.gemm_dx_prep_supported(64, 16)
    dP = ops.split_bf16(torch.randn((M, K), generator=g, device=dev), True)
    WT = ops.split_bf16(torch.randn((N, K), generator=g, device=dev) / K ** 0.5, True)
    y = torch.randn((M, N), generator=g, device=dev)
    bias = torch.randn((M,), generator=g, device=dev)
    drop = ops.DropMask.feature_mask(torch.rand((M, N), generator=g, device=dev) >= 0.3, 0.3) if with_drop else None
    if with_drop:  # what the forward would have stored: post-dropout rows
        cols = torch.arange(N, device=dev)
        keep = ((drop.bits[:, cols // 32] >> (cols % 32)) & 1).bool()
        y = torch.where(keep, y, torch.zeros_like(y))
    G, t, hsum = ops.gemm_dx_prep(dP, WT, M, N, K, y, bias, H, F, apply_elu=True, feat_drop=drop)
    dX = ops.gemm(dP, False, WT, False, M, N, K)
    G2, t2, hsum2 = ops.edge_bwd_prep(dX, y, bias, H, F, apply_elu=True, feat_drop=drop)
    assert rel_err(G.cpu().numpy(), G2.cpu().numpy()) < 1e-6
    assert rel_err(t.cpu().numpy(), t2.cpu().numpy()) < 1e-5
    assert rel_err(hsum.cpu().numpy(), hsum2.cpu().numpy()) < 1e-5


def test_fused_gelu_layernorm_matches_torch(dev):
    """ProjectionHead hidden block (reference core/model/projection.py:56-62): LayerNorm(GELU(h)) fused, forward and
    every gradient against torch's fp64 ops; and the head as a whole against the ATen composition of the same modules."""
    g = torch.Generator().manual_seed(4)
    M, D = 777, 1600
    h = torch.randn(M, D, generator=g) * 2
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    dy = torch.randn(M, D, generator=g)
    hr, gr, br = (t.double().clone().requires_grad_(True) for t in (h, gamma, beta))
    ref = torch.nn.functional.layer_norm(torch.nn.functional.gelu(hr), (D,), gr, br, 1e-5)
    ref.backward(dy.double())
    hx, gx, bx = (t.to(dev).requires_grad_(True) for t in (h, gamma, beta))
    y = RF.gelu_layernorm(hx, gx, bx, 1e-5)
    y.backward(dy.to(dev))
    assert rel_err(y.detach().cpu().numpy(), ref.detach().numpy()) < 1e-5
    assert rel_err(hx.grad.cpu().numpy(), hr.grad.numpy()) < 1e-5
    assert rel_err(gx.grad.cpu().numpy(), gr.grad.numpy()) < 1e-5
    assert rel_err(bx.grad.cpu().numpy(), br.grad.numpy()) < 1e-5
    torch.manual_seed(0)
    head = R.ProjectionHead(64, 48, num_layers=3, hidden_dim=80).to(dev)
    x = torch.randn(33, 64, device=dev)
    want = x
    for mod in head.net:  # the same modules, composed with ATen ops
        want = torch.nn.functional.linear(want, mod.weight) if isinstance(mod, torch.nn.Linear) else mod(want)
    assert rel_err(head(x).detach().cpu().numpy(), want.detach().cpu().numpy()) < 1e-4


def test_second_generation_by_source_kernel_equals_the_first(dev, monkeypatch):
    """csrc/edge_bwd_src2.cu (opt-in, RELGAT_SRC_V2=1) against the default by-source kernel on the training path:
    same dP planes and dS columns (rounding-level differences: the attention weights are rebuilt in a pre-pass)."""
    from relgat_projector_b200.graph import GraphIndex
    gen = torch.Generator(device=dev).manual_seed(8)
    n, e, r, h, f = 3000, 20000, 11, 4, 40
    ei = torch.randint(0, n, (2, e), generator=gen, device=dev)
    ei[1, :1500] = 7  # a destination / source hub exercises the split-segment path
    ei[0, 1500:2200] = 9
    et = torch.randint(0, r, (e,), generator=gen, device=dev)
    g = GraphIndex(ei, et, n, r)
    P = torch.randn((n, h * f), generator=gen, device=dev)
    A = torch.randn((h, r, f), generator=gen, device=dev) * 0.2
    beta = torch.randn((r,), generator=gen, device=dev) * 0.1
    dY = torch.randn((n, h * f), generator=gen, device=dev)
    out, _, _, z, minv, bias = ops.edge_fwd(P, A, beta, g, h, f)
    G, t, _ = ops.edge_bwd_prep(dY, out, bias, h, f, apply_elu=False)
    res = []
    for v2 in (False, True):
        monkeypatch.setattr(ops, "SRC_V2", v2)
        _, (hi, lo), _ = ops.edge_bwd_src(P, G, A, z, minv, t, g, h, f, want_fp32=False, want_planes=True, want_ds=True)
        res.append(hi.float() + lo.float())
    assert rel_err(res[1].cpu().numpy(), res[0].cpu().numpy()) < 1e-5


# ------------------------------------------------------------------------------------------------------
# exact-zero rows of the output gradient (functional.SPARSE_BWD)
# ------------------------------------------------------------------------------------------------------
def _bits_to_rows(bits: torch.Tensor, n: int) -> np.ndarray:
    w = bits.cpu().numpy().view(np.uint32)
    return np.nonzero((w[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1)[0]


def test_row_bitmaps_mark_rows_and_their_sources(dev):
    gen = torch.Generator(device=dev).manual_seed(3)
    n, e, r = 5000, 21000, 7
    ei = torch.randint(0, n, (2, e), generator=gen, device=dev)
    et = torch.randint(0, r, (e,), generator=gen, device=dev)
    g = GraphIndex(ei, et, n, r)
    ids = torch.randint(0, n, (300,), generator=gen, device=dev)
    bits = ops.mark_rows(ids, n)
    assert np.array_equal(_bits_to_rows(bits, n), np.unique(ids.cpu().numpy()))
    src_bits = ops.mark_sources(bits, g)
    marked = torch.zeros(n, dtype=torch.bool, device=dev)
    marked[ids] = True
    want = torch.unique(ei[0][marked[ei[1]]]).cpu().numpy()
    assert np.array_equal(_bits_to_rows(src_bits, n), want)
    # empty id list: nothing marked; out-of-range ids are ignored rather than written out of bounds
    assert _bits_to_rows(ops.mark_rows(ids[:0], n), n).size == 0
    assert np.array_equal(_bits_to_rows(ops.mark_rows(torch.tensor([-1, 5, n, n + 77], device=dev), n), n), [5])


@pytest.mark.parametrize("storage", ["fp32", "bf16"])
def test_by_source_pass_with_zero_row_hint_equals_the_full_pass(dev, storage):
    """Edges into rows of G that are exact zeros contribute exact zeros: skipping them changes no bit of [dP | dS]."""
    gen = torch.Generator(device=dev).manual_seed(9)
    n, e, r, h, f = 3000, 20000, 11, 4, 40
    ei = torch.randint(0, n, (2, e), generator=gen, device=dev)
    ei[1, :1500] = 7   # destination hub (in the non-zero set below)
    ei[0, 1500:2200] = 9  # source hub: split-segment path with skipped edges inside its parts
    et = torch.randint(0, r, (e,), generator=gen, device=dev)
    g = GraphIndex(ei, et, n, r)
    dt = torch.float32 if storage == "fp32" else torch.bfloat16
    P = torch.randn((n, h * f), generator=gen, device=dev)
    A = torch.randn((h, r, f), generator=gen, device=dev) * 0.2
    beta = torch.randn((r,), generator=gen, device=dev) * 0.1
    rows = torch.unique(torch.cat([torch.randint(0, n, (150,), generator=gen, device=dev),
                                   torch.tensor([7, 0, n - 1], device=dev)]))
    dY = torch.zeros((n, h * f), device=dev)
    dY[rows] = torch.randn((rows.numel(), h * f), generator=gen, device=dev)
    out, _, _, z, minv, bias = ops.edge_fwd(P.to(dt), A, beta, g, h, f)
    G, t, _ = ops.edge_bwd_prep(dY, out, bias, h, f, apply_elu=False, g_bf16=storage == "bf16")
    res = []
    for hint in (None, ops.mark_rows(rows, n)):
        _, (hi, lo), _ = ops.edge_bwd_src(P.to(dt), G, A, z, minv, t, g, h, f, want_fp32=False, want_planes=True,
                                          want_ds=True, dst_nz=hint)
        res.append((hi.clone(), lo.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert float(res[0][0].float().abs().sum()) > 0
    # the hint is refused where it would leave dz unwritten, and when it is not a bitmap of the graph's rows
    with pytest.raises(ValueError):
        ops.edge_bwd_src(P.to(dt), G, A, z, minv, t, g, h, f, want_ds=False, dst_nz=ops.mark_rows(rows, n))
    with pytest.raises(ValueError):
        ops.edge_bwd_src(P.to(dt), G, A, z, minv, t, g, h, f, want_ds=True, dst_nz=torch.zeros(3, dtype=torch.int32, device=dev))


@pytest.mark.parametrize("compact", [False, True])
@pytest.mark.parametrize("name", ["f200_fp32", "transe_proj_fp32", "tiny_fp64", "adversarial_fp32"])
def test_training_step_with_and_without_zero_row_skipping(dev, name, compact, monkeypatch):
    """Skipping the edges into exact-zero rows changes no bit of the gradients; compacting the backward to the rows that
    can be non-zero (fewer rows in the weight-gradient GEMMs' sums) changes only their rounding."""
    c = Case(name)
    res = []
    monkeypatch.setattr(RF, "COMPACT_BWD", compact)
    for sparse in (True, False):
        monkeypatch.setattr(RF, "SPARSE_BWD", sparse)
        m = _load_model(c, dev)
        m.train()
        src, rel, dst = (c.t(k).to(dev) for k in ("src_ids", "rel_ids", "dst_ids"))
        scores = m(src, rel, dst)
        scores = scores[0] if isinstance(scores, tuple) else scores  # (scores, transform rows, ...) with a projection
        (scores * torch.linspace(-1, 1, scores.numel(), device=dev)).sum().backward()
        res.append({n_: p.grad.clone() for n_, p in m.named_parameters() if p.grad is not None})
    assert res[0].keys() == res[1].keys() and len(res[0]) > 0
    for n_ in res[0]:
        if compact:
            assert rel_err(res[0][n_].cpu().numpy(), res[1][n_].cpu().numpy()) < 2e-5, n_
        else:
            assert torch.equal(res[0][n_], res[1][n_]), n_


# ------------------------------------------------------------------------------------------------------
# receptive-field blocks
# ------------------------------------------------------------------------------------------------------
def test_blocks_hold_every_in_edge_of_their_destinations_in_graph_order(dev):
    from relgat_projector_b200.blocks import build_blocks
    gen = torch.Generator(device=dev).manual_seed(5)
    n, e, r = 4000, 18000, 9
    ei = torch.randint(0, n, (2, e), generator=gen, device=dev)
    ei[1, :700] = 11  # a hub with more in-edges than a split segment holds
    et = torch.randint(0, r, (e,), generator=gen, device=dev)
    full = GraphIndex(ei, et, n, r)
    ids = torch.randint(0, n, (200,), generator=gen, device=dev)
    ids[:3] = torch.tensor([11, 11, 0], device=dev)
    blk = build_blocks(full, ids, 2)
    torch.cuda.synchronize()
    rowptr, csr_src, csr_rel = (getattr(full, k).cpu().numpy() for k in ("rowptr", "csr_src", "csr_rel"))
    rows = blk.input_rows.cpu().numpy()            # D_0
    D = np.unique(ids.cpu().numpy())               # D_2
    assert np.array_equal(D[blk.out_pos.cpu().numpy()], ids.cpu().numpy())
    for g in reversed(blk.graphs):                 # walk down from the last block
        assert g.N == D.size
        bp, bs, br = (getattr(g, k).cpu().numpy() for k in ("rowptr", "csr_src", "csr_rel"))
        srcs = np.unique(np.concatenate([csr_src[rowptr[j]:rowptr[j + 1]] for j in D] + [np.zeros(0, np.int32)]))
        assert g.N_src == srcs.size
        for k, j in enumerate(D):                  # same edges, same order, sources relabelled into the sorted set
            assert np.array_equal(srcs[bs[bp[k]:bp[k + 1]]], csr_src[rowptr[j]:rowptr[j + 1]])
            assert np.array_equal(br[bp[k]:bp[k + 1]], csr_rel[rowptr[j]:rowptr[j + 1]])
        D = srcs
    assert np.array_equal(rows, D)
    assert blk.n_edges == sum(g.E for g in blk.graphs)


@pytest.mark.parametrize("mode", ["masked", "blocks"])
@pytest.mark.parametrize("name", MODEL_CASES)
def test_receptive_field_step_equals_the_full_graph_step(dev, name, mode):
    """forward on the batch's blocks: the same scores (bit for bit — every destination keeps all its in-edges in order)
    and the same parameter gradients (the weight-gradient GEMMs sum over fewer, reordered rows: rounding only)."""
    c = Case(name)
    res = []
    for rf in (False, True):
        m = _load_model(c, dev)
        m.receptive_field, m.receptive_field_mode = rf, mode
        src, rel, dst = (c.t(k).to(dev) for k in ("src_ids", "rel_ids", "dst_ids"))
        scores = m(src, rel, dst)[0]
        (scores * torch.linspace(-1, 1, scores.numel(), device=dev)).sum().backward()
        res.append((scores.detach().clone(), {n_: p.grad.clone() for n_, p in m.named_parameters() if p.grad is not None}))
        if rf:
            assert m.last_block_edges is not None and m.last_block_edges <= c.layers * c.e
    assert torch.equal(res[0][0], res[1][0])
    assert res[0][1].keys() == res[1][1].keys() and len(res[0][1]) > 0
    for n_ in res[0][1]:
        assert rel_err(res[1][1][n_].cpu().numpy(), res[0][1][n_].cpu().numpy()) < 2e-5, n_


@pytest.mark.parametrize("mode", ["masked", "blocks"])
def test_receptive_field_step_with_dropout_and_losses_runs_and_is_finite(dev, mode):
    c = Case("transe_proj_fp32")
    m = _load_model(c, dev)
    m.receptive_field, m.receptive_field_mode = True, mode
    for lyr in m._layers():
        lyr.dropout.p, lyr.rel_attn_drop.p = 0.3, 0.1
    torch.manual_seed(0)
    rank = L.RelGATLoss("self_adversarial_loss", 1.0, None, None, {})
    multi = L.MultiObjectiveRelLoss(relgat_loss=rank, run_config={})
    src, rel, dst = (c.t(k).to(dev) for k in ("src_ids", "rel_ids", "dst_ids"))
    _, _, loss, *_ = L.calculate_loss(m, src, rel, dst, c.b, rank, multi)
    loss.backward()
    assert torch.isfinite(loss)
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)


@pytest.mark.parametrize("shape", ["uniform", "hubs", "empty_segments", "single", "none"])
def test_device_built_work_tables_equal_the_host_formulation(dev, shape):
    """csrc/stream_chunks.cu against graph.py::StreamChunks (torch ops, used once per full graph): same tables."""
    from relgat_projector_b200.graph import StreamChunks
    gen = torch.Generator().manual_seed(11)
    if shape == "uniform":
        deg = torch.randint(0, 12, (5000,), generator=gen)
    elif shape == "hubs":
        deg = torch.randint(0, 9, (3000,), generator=gen)
        deg[[0, 17, 18, 1500, 2999]] = torch.tensor([513, 2000, 700, 5121, 512])
    elif shape == "empty_segments":
        deg = torch.zeros(300, dtype=torch.int64)
        deg[[5, 250]] = torch.tensor([40, 3])
    elif shape == "single":
        deg = torch.tensor([7])
    else:
        deg = torch.zeros(0, dtype=torch.int64)
    ptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(deg, 0)]).to(torch.int32).to(dev)
    want = StreamChunks(ptr)
    got = StreamChunks.launch(ptr, int(deg.sum()))
    got.finish(got.counts.cpu().numpy())
    assert (got.n_chunks, got.n_parts, got.n_long) == (want.n_chunks, want.n_parts, want.n_long)
    for k in ("chunks", "parts", "long_node", "long_part_ptr"):
        assert torch.equal(getattr(got, k), getattr(want, k)), k


@pytest.mark.parametrize("mode", ["masked", "blocks"])
def test_config1_full_size_receptive_field_step_vs_full_graph_step(dev, mode):
    """Config 1 (10 k nodes / 45 k edges / 1024-d / 2 layers / 4x200, B = 256, K = 4) through calculate_loss: the step on
    the batch's blocks against the full-graph step (itself checked against the fp64 oracle above)."""
    cfg = S.CONFIGS["c1"]
    kg = S.tensor_kg(cfg["N"], cfg["T"], cfg["R"], cfg["D_in"], seed=3, device=dev)
    gen = torch.Generator().manual_seed(5)
    src, rel, dst = (t.to(dev) for t in S.sample_batch(kg.train_triples.cpu(), cfg["N"], cfg["B"], cfg["K"], gen))
    rank = L.RelGATLoss("margin", None, 1.0, None, {})
    res = []
    for rf in (False, True):
        torch.manual_seed(11)
        m = R.RelGATModel(kg.node_emb, kg.edge_index, kg.edge_type, num_rel=cfg["R"], scorer_type=cfg["scorer"],
                          gat_out_dim=cfg["F"], gat_heads=cfg["H"], dropout=0.0, gat_num_layers=cfg["L"]).to(dev).train()
        m.receptive_field, m.receptive_field_mode = rf, mode
        _, _, loss, *_ = L.calculate_loss(m, src, rel, dst, cfg["B"], rank, None)
        loss.backward()
        res.append((float(loss), {n_: p.grad.clone() for n_, p in m.named_parameters() if p.grad is not None}))
        if rf:
            assert 0 < m.last_block_edges < cfg["L"] * kg.edge_index.size(1)
    assert res[0][0] == res[1][0]  # every batch row is computed from the same terms in the same order
    for n_ in res[0][1]:
        assert rel_err(res[1][1][n_].cpu().numpy(), res[0][1][n_].cpu().numpy()) < 2e-5, n_


@pytest.mark.parametrize("mode", ["masked", "blocks"])
def test_receptive_field_edge_cases_isolated_nodes_and_repeated_ids(dev, mode):
    """Batch nodes without in-edges (empty blocks all the way down), a batch naming one node many times, and a one-layer
    model: the block path must agree with the full-graph path."""
    gen = torch.Generator().manual_seed(2)
    n, r, d_in = 400, 5, 32
    ei = torch.randint(0, 300, (2, 1500), generator=gen)      # nodes 300..399 are isolated
    et = torch.randint(0, r, (1500,), generator=gen)
    x0 = torch.randn(n, d_in, generator=gen)
    for layers, ids in ((2, torch.tensor([350, 399, 350])),             # isolated only: every block has E = 0
                        (2, torch.tensor([7, 7, 7, 350, 7])),           # repeats + an isolated node
                        (1, torch.randint(0, n, (64,), generator=gen))):
        res = []
        for rf in (False, True):
            torch.manual_seed(4)
            m = R.RelGATModel(x0.to(dev), ei.to(dev), et.to(dev), num_rel=r, gat_out_dim=8, gat_heads=2, dropout=0.0,
                              gat_num_layers=layers).to(dev).train()
            m.receptive_field, m.receptive_field_mode = rf, mode
            rows = m.batch_rows(ids.to(dev))
            (rows * torch.linspace(-1, 1, rows.numel(), device=dev).view_as(rows)).sum().backward()
            res.append((rows.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
        assert torch.equal(res[0][0], res[1][0])
        assert res[0][1].keys() == res[1][1].keys()
        for k in res[0][1]:
            assert rel_err(res[1][1][k].cpu().numpy(), res[0][1][k].cpu().numpy()) < 2e-5, (layers, k)


def test_bitmap_ranks_numbering(dev):
    gen = torch.Generator(device=dev).manual_seed(4)
    for n in (1, 31, 32, 33, 5000):
        ids = torch.randint(0, n, (max(1, n // 3),), generator=gen, device=dev)
        bits = ops.mark_rows(ids, n)
        rank, lst, cnt = ops.bitmap_ranks(bits, n)
        want = np.unique(ids.cpu().numpy())
        assert int(cnt) == want.size
        assert np.array_equal(lst[:want.size].cpu().numpy(), want)
        r = rank.cpu().numpy()
        assert np.array_equal(np.nonzero(r >= 0)[0], want) and np.array_equal(r[want], np.arange(want.size))
    rank, lst, cnt = ops.bitmap_ranks(ops.row_bitmap(100, dev), 100)
    assert int(cnt) == 0 and bool((rank == -1).all())


def test_compacted_by_source_pass_equals_rows_of_the_full_pass(dev):
    """src_rows: the output holds exactly the rows of the sources with an edge into a marked row, in ascending source
    order, bit-identical to those rows of the uncompacted output; every other row of the full output is zero."""
    gen = torch.Generator(device=dev).manual_seed(10)
    n, e, r, h, f = 3000, 20000, 11, 4, 40
    ei = torch.randint(0, n, (2, e), generator=gen, device=dev)
    ei[1, :1500] = 7
    ei[0, 1500:2200] = 9     # split source, reaches marked rows
    ei[0][ei[0] == 10] = 11  # source 10 gets exactly the 700 edges below
    ei[0, 2200:2900] = 10    # split source ...
    ei[1, 2200:2900] = torch.randint(100, 200, (700,), generator=gen, device=dev)  # ... into unmarked rows only
    et = torch.randint(0, r, (e,), generator=gen, device=dev)
    g = GraphIndex(ei, et, n, r)
    P = torch.randn((n, h * f), generator=gen, device=dev)
    A = torch.randn((h, r, f), generator=gen, device=dev) * 0.2
    beta = torch.randn((r,), generator=gen, device=dev) * 0.1
    rows = torch.unique(torch.cat([torch.randint(200, n, (120,), generator=gen, device=dev), torch.tensor([7], device=dev)]))
    dY = torch.zeros((n, h * f), device=dev)
    dY[rows] = torch.randn((rows.numel(), h * f), generator=gen, device=dev)
    out, _, _, z, minv, bias = ops.edge_fwd(P, A, beta, g, h, f)
    G, t, _ = ops.edge_bwd_prep(dY, out, bias, h, f, apply_elu=False)
    bits = ops.mark_rows(rows, n)
    _, (hi, lo), _ = ops.edge_bwd_src(P, G, A, z, minv, t, g, h, f, want_fp32=False, want_planes=True, want_ds=True, dst_nz=bits)
    rank, lst, cnt = ops.bitmap_ranks(ops.mark_sources(bits, g), n)
    n_s = int(cnt)
    _, (chi, clo), _ = ops.edge_bwd_src(P, G, A, z, minv, t, g, h, f, want_fp32=False, want_planes=True, want_ds=True,
                                        dst_nz=bits, src_rows=(rank, n_s))
    keep = lst[:n_s]
    assert 0 < n_s < n and chi.size(0) == n_s
    assert torch.equal(chi, hi[keep]) and torch.equal(clo, lo[keep])
    other = torch.ones(n, dtype=torch.bool, device=dev)
    other[keep] = False
    assert float(hi[other].float().abs().sum()) == 0.0 and bool(other[10])
    # compact prep: the gradient rows of the listed nodes only -> rows of a zero table
    Gt = torch.zeros_like(dY)
    G2, t2, h2 = ops.edge_bwd_prep(dY[rows].contiguous(), out, bias, h, f, apply_elu=True, G_out=Gt, compact_rows=rows)
    G1, t1, h1 = ops.edge_bwd_prep(dY, out, bias, h, f, apply_elu=True)
    assert torch.equal(G2, G1) and torch.equal(t2, t1) and torch.equal(h2, h1)


def test_masked_receptive_field_step_uses_the_same_dropout_masks_as_the_full_graph_step(dev):
    """The masked mode keeps the full graph's row and edge numbering, so a given seed draws the same dropout masks:
    with dropout active the batch rows still equal the full-graph step's bit for bit."""
    c = Case("transe_proj_fp32")
    res = []
    for rf in (False, True):
        m = _load_model(c, dev)
        m.receptive_field, m.receptive_field_mode = rf, "masked"
        for lyr in m._layers():
            lyr.dropout.p, lyr.rel_attn_drop.p = 0.3, 0.2
        torch.manual_seed(21)
        ids = torch.cat([c.t("src_ids"), c.t("dst_ids")]).to(dev)
        rows = m._stack_output(gather_ids=ids)
        rows.square().sum().backward()
        res.append((rows.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    assert torch.equal(res[0][0], res[1][0])
    for k in res[0][1]:
        assert rel_err(res[1][1][k].cpu().numpy(), res[0][1][k].cpu().numpy()) < 2e-5, k


def test_work_table_over_a_list_of_destinations(dev):
    from relgat_projector_b200.graph import LONG_SEGMENT, PART_EDGES, StreamChunks
    gen = torch.Generator().manual_seed(13)
    deg = torch.randint(0, 9, (2000,), generator=gen)
    deg[[3, 900]] = torch.tensor([LONG_SEGMENT + 1, 3 * PART_EDGES + 5])
    ptr = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(deg, 0)]).to(torch.int32).to(dev)
    rows = torch.tensor([0, 3, 4, 17, 900, 1999], dtype=torch.int64)
    padded = torch.cat([rows, torch.full((50,), 12345678, dtype=torch.int64)]).to(dev)  # entries beyond the count: ignored
    ck = StreamChunks.launch_for_rows(ptr, int(deg.sum()), padded, torch.tensor([rows.numel()], dtype=torch.int32, device=dev))
    ck.finish(ck.counts.cpu().numpy())
    p = ptr.cpu().numpy()
    parts, chunks = [], []
    for j in (3, 900):
        for lo in range(p[j], p[j + 1], PART_EDGES):
            chunks.append([j, 1, len(parts), 0])
            parts.append([lo, min(lo + PART_EDGES, p[j + 1])])
    chunks += [[j, 1, -1, 0] for j in (0, 4, 17, 1999)]
    assert (ck.n_chunks, ck.n_parts, ck.n_long) == (len(chunks), len(parts), 2)
    assert ck.chunks.cpu().tolist() == chunks and ck.parts.cpu().tolist() == parts
    assert ck.long_node.cpu().tolist() == [3, 900] and ck.long_part_ptr.cpu().tolist() == [0, 3, 7]
    assert ck.n_edges == int(deg[rows].sum())


@pytest.mark.parametrize("mode", [None, "masked", "blocks"])
def test_batch_rows_with_an_empty_id_list(dev, mode):
    """No row requested: an empty result, and a backward that yields all-zero parameter gradients (full-graph step,
    compacted backward and both receptive-field modes)."""
    c = Case("tiny_fp64")
    m = _load_model(c, dev)
    if mode:
        m.receptive_field, m.receptive_field_mode = True, mode
    rows = m.batch_rows(torch.empty(0, dtype=torch.int64, device=dev))
    assert tuple(rows.shape) == (0, c.h * c.f)
    rows.sum().backward()
    grads = [p.grad for n_, p in m.named_parameters() if n_.startswith("gat_layer") and p.grad is not None]
    assert grads and all(float(g_.abs().sum()) == 0.0 for g_ in grads)
