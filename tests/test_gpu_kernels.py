"""GPU parity tests, kernel level: every C-ABI entry point against the CPU oracle on the same
seeded inputs.  Integer structures bit-exact; floating point within 1e-4 relative
(BASELINE.json north_star), tolerance stated per assert."""
import numpy as np
import pytest
import torch

from oracle import relgat_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4   # north_star: "within 1e-4 relative in fp32"
GEMM_SPLIT_TOL = 2e-5   # bf16 hi/lo split GEMM vs an fp64 product of the fp32 inputs
GEMM_BF16_TOL = 1e-5    # plain bf16 GEMM vs an fp64 product of the bf16-rounded inputs


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _graph(rng, n, e, r, isolated=0, hub=0):
    src = rng.integers(0, n, size=e).astype(np.int64)
    dst = rng.integers(0, max(1, n - isolated), size=e).astype(np.int64)
    rel = rng.integers(0, r, size=e).astype(np.int64)
    if hub:
        dst[-hub:] = 1
        src[-hub // 2:] = 2
    return src, dst, rel


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,e,r,isolated,hub", [(37, 500, 5, 4, 0), (1000, 20000, 50, 100, 3000), (5, 0, 2, 0, 0),
                                               (1, 7, 1, 0, 0), (70000, 300000, 200, 0, 0)])
def test_graph_index_bit_exact(dev, n, e, r, isolated, hub):
    from relgat_projector_b200.graph import GraphIndex
    rng = np.random.default_rng(n + e)
    src, dst, rel = _graph(rng, n, e, r, isolated, hub)
    ei = torch.from_numpy(np.stack([src, dst])).to(dev)
    et = torch.from_numpy(rel).to(dev)
    g = GraphIndex(ei, et, n, r).as_numpy()
    ref = O.graph_index_np(src, dst, rel, n, r)
    for k, v in ref.items():
        assert g[k].dtype == np.int32 and np.array_equal(g[k], v), k


def test_graph_index_rejects_bad_input(dev):
    from relgat_projector_b200.graph import GraphIndex
    ei = torch.tensor([[0, 5], [1, 2]], device=dev)
    with pytest.raises(IndexError):
        GraphIndex(ei, torch.tensor([0, 0], device=dev), 3, 1)
    with pytest.raises(TypeError):
        GraphIndex(ei.int(), torch.tensor([0, 0], device=dev), 9, 1)


# ------------------------------------------------------------------------------------------
GEMM_SHAPES = [
    # M, N, K
    (128, 64, 64), (300, 200, 136), (1000, 800, 1024), (257, 48, 72), (130, 1024, 200), (4096, 160, 64),
    # CTA-pair tiles (256 rows): a second CTA whose rows are all / partly outside M, odd tile counts, 16-column output
    (129, 16, 64), (513, 1000, 128), (1024, 256, 192), (40000, 800, 264),
]


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("split", [True, False])
def test_gemm_all_layouts(dev, a_mn, b_mn, split):
    from relgat_projector_b200 import ops
    g = torch.Generator(device=dev).manual_seed(11)
    for (M, N, K) in GEMM_SHAPES:
        if (a_mn and M % 8) or (b_mn and N % 8) or ((not a_mn or not b_mn) and K % 8):
            continue  # TMA needs 16-byte row strides
        A = torch.randn((M, K), generator=g, device=dev)
        B = torch.randn((N, K), generator=g, device=dev)
        a_store = A.t().contiguous() if a_mn else A
        b_store = B.t().contiguous() if b_mn else B
        ap, bp = ops.split_bf16(a_store, split), ops.split_bf16(b_store, split)
        D = ops.gemm(ap, a_mn, bp, b_mn, M, N, K)
        if split:
            ref = A.double() @ B.double().t()
            tol = GEMM_SPLIT_TOL
        else:
            ref = A.bfloat16().double() @ B.bfloat16().double().t()
            tol = GEMM_BF16_TOL
        err = rel_err(D.cpu().numpy(), ref.cpu().numpy())
        assert err < tol, (M, N, K, a_mn, b_mn, split, err)


@pytest.mark.parametrize("M,N,K", [(1000, 800, 1024), (700, 208, 136), (3000, 1000, 200)])
def test_gemm_bf16_output_tiles_do_not_overlap(dev, M, N, K):
    """bf16 output (the bf16 feature-storage mode): N tiles whose width is not a multiple of the 32-column store box
    must not spill into their neighbours."""
    from relgat_projector_b200 import ops
    g = torch.Generator(device=dev).manual_seed(21)
    A = torch.randn((M, K), generator=g, device=dev)
    B = torch.randn((N, K), generator=g, device=dev)
    for split in (False, True):
        D = ops.gemm(ops.split_bf16(A, split), False, ops.split_bf16(B, split), False, M, N, K, out_dtype=torch.bfloat16)
        ref = (A.double() @ B.double().t()) if split else (A.bfloat16().double() @ B.bfloat16().double().t())
        assert rel_err(D.float().cpu().numpy(), ref.cpu().numpy()) < 1e-2  # bf16 rounding of the output


def test_split_planes_reconstruct_fp32(dev):
    from relgat_projector_b200 import ops
    x = torch.randn(1000, 77, device=dev) * 37.0
    hi, lo = ops.split_bf16(x)
    rec = hi.float() + lo.float()
    assert rel_err(rec.cpu().numpy(), x.cpu().numpy()) < 2 ** -15
    assert torch.equal(hi, x.bfloat16())


@pytest.mark.parametrize("M,N,K,splits", [(800, 1024, 5000, 5), (160, 64, 1000, 16), (128, 128, 64, 4)])
def test_gemm_split_k_mn_major(dev, M, N, K, splits):
    """The dW shape: both operands MN-major, long reduction over nodes, ordered split-K."""
    from relgat_projector_b200 import ops
    g = torch.Generator(device=dev).manual_seed(5)
    dP = torch.randn((K, M), generator=g, device=dev)
    X = torch.randn((K, N), generator=g, device=dev)
    D = ops.gemm(ops.split_bf16(dP), True, ops.split_bf16(X), True, M, N, K, splits_k=splits)
    ref = dP.double().t() @ X.double()
    assert rel_err(D.cpu().numpy(), ref.cpu().numpy()) < GEMM_SPLIT_TOL
    D2 = ops.gemm(ops.split_bf16(dP), True, ops.split_bf16(X), True, M, N, K, splits_k=splits)
    assert torch.equal(D, D2)  # ordered reduction: bitwise reproducible


# ------------------------------------------------------------------------------------------
LAYER_SHAPES = [
    # n, e, r, H, F, isolated, hub
    (50, 400, 7, 3, 8, 5, 0),        # the tiny case of SURVEY §8(c); H=3 -> one head per warp
    (300, 2500, 11, 4, 200, 20, 0),  # head width of the named configs (C = 800)
    (200, 1500, 5, 8, 200, 0, 0),    # config-3 heads (two warps per node)
    (120, 900, 6, 2, 10, 0, 0),      # F % 4 != 0 -> scalar path
    (400, 6000, 9, 16, 128, 30, 2000),  # the repo's own training script shape, with a hub node
    (64, 0, 3, 4, 16, 0, 0),         # no edges at all
]


@pytest.mark.parametrize("n,e,r,H,F,isolated,hub", LAYER_SHAPES)
def test_edge_forward_and_backward_vs_closed_form(dev, n, e, r, H, F, isolated, hub):
    from relgat_projector_b200 import ops
    from relgat_projector_b200.graph import GraphIndex
    rng = np.random.default_rng(e + H)
    src, dst, rel = _graph(rng, n, e, r, isolated, hub)
    P = rng.standard_normal((n, H, F)).astype(np.float32)
    A = (rng.standard_normal((H, r, F)) / np.sqrt(F)).astype(np.float32)
    beta = (rng.standard_normal(r) * 0.1).astype(np.float32)
    gi = O.graph_index_np(src, dst, rel, n, r)
    out_ref, z_ref, alpha_ref, bias_ref = O.layer_forward_closed(P, A, beta, gi)

    g = GraphIndex(torch.from_numpy(np.stack([src, dst])).to(dev), torch.from_numpy(rel).to(dev), n, r)
    Pd = torch.from_numpy(P.reshape(n, H * F)).to(dev)
    Ad, bd = torch.from_numpy(A).to(dev), torch.from_numpy(beta).to(dev)
    out, act, alpha, z, minv, bias = ops.edge_fwd(Pd, Ad, bd, g, H, F, want_act=True, apply_elu=True, want_alpha=True)
    assert rel_err(out.cpu().numpy(), out_ref.reshape(n, -1)) < FP32_TOL
    if e:
        assert rel_err(z.cpu().numpy(), z_ref) < FP32_TOL
        assert rel_err(alpha.cpu().numpy(), alpha_ref) < FP32_TOL
        assert np.allclose(alpha.cpu().numpy().sum(0), alpha_ref.sum(0), rtol=1e-4)
    assert rel_err(bias.cpu().numpy(), bias_ref) < FP32_TOL
    deg = np.diff(gi["rowptr"])
    assert np.all(out.cpu().numpy()[deg == 0] == 0.0)  # isolated destinations are exactly zero
    # fused ELU + bf16 split epilogue
    elu = np.where(out_ref > 0, out_ref, np.expm1(np.minimum(out_ref, 0))).reshape(n, -1)
    rec = act[0].float() + act[1].float()
    assert rel_err(rec.cpu().numpy(), elu) < FP32_TOL

    # backward
    Gn = rng.standard_normal((n, H, F)).astype(np.float32)
    dP_ref, dA_ref, dbeta_ref, dz_ref = O.layer_backward_closed(Gn, P, A, gi, z_ref, alpha_ref)
    Gd = torch.from_numpy(Gn.reshape(n, -1)).to(dev)
    G, t, hsum = ops.edge_bwd_prep(Gd, out, bias, H, F, apply_elu=False)
    assert torch.equal(G, Gd)
    dP, planes, dz = ops.edge_bwd_src(Pd, G, Ad, z, minv, t, g, H, F, want_fp32=True, want_planes=True)
    dA, dbeta = ops.edge_bwd_rel(Pd, dz, hsum, g, H, F)
    assert rel_err(dP.cpu().numpy(), dP_ref.reshape(n, -1)) < FP32_TOL
    if e:
        assert rel_err(dz.cpu().numpy(), dz_ref) < FP32_TOL
    assert rel_err(dA.cpu().numpy(), dA_ref) < FP32_TOL
    assert rel_err(dbeta.cpu().numpy(), dbeta_ref) < FP32_TOL
    assert rel_err((planes[0].float() + planes[1].float()).cpu().numpy(), dP_ref.reshape(n, -1)) < FP32_TOL
    # determinism: identical bits on a second run
    out2, _, alpha2, _, _, _ = ops.edge_fwd(Pd, Ad, bd, g, H, F, want_alpha=True)
    dP2, _, dz2 = ops.edge_bwd_src(Pd, G, Ad, z, minv, t, g, H, F)
    dA2, dbeta2 = ops.edge_bwd_rel(Pd, dz, hsum, g, H, F)
    assert torch.equal(out, out2) and torch.equal(alpha, alpha2) and torch.equal(dP, dP2)
    assert torch.equal(dA, dA2) and torch.equal(dbeta, dbeta2)


@pytest.mark.parametrize("n,e,r,H,F,isolated,hub", [s_ for s_ in LAYER_SHAPES if s_[4] % 4 == 0])
def test_bwd_src_third_generation_matches_first(dev, n, e, r, H, F, isolated, hub):
    """relgat_layer_bwd_src3 (bulk-copy ring, no attention-vector term): its dS columns equal the first generation's,
    its dPa rows equal the first generation's dP minus dS·A (fp32 rows and bf16 planes), bitwise reproducible."""
    from relgat_projector_b200 import ops
    from relgat_projector_b200.graph import GraphIndex
    rng = np.random.default_rng(e + H + 1)
    src, dst, rel = _graph(rng, n, e, r, isolated, hub)
    g = GraphIndex(torch.from_numpy(np.stack([src, dst])).to(dev), torch.from_numpy(rel).to(dev), n, r)
    Pd = torch.from_numpy(rng.standard_normal((n, H * F)).astype(np.float32)).to(dev)
    Ad = torch.from_numpy((rng.standard_normal((H, r, F)) / np.sqrt(F)).astype(np.float32)).to(dev)
    bd = torch.from_numpy((rng.standard_normal(r) * 0.1).astype(np.float32)).to(dev)
    out, _, _, z, minv, bias = ops.edge_fwd(Pd, Ad, bd, g, H, F)
    Gd = torch.from_numpy(rng.standard_normal((n, H * F)).astype(np.float32)).to(dev)
    G, t, _ = ops.edge_bwd_prep(Gd, out, bias, H, F, apply_elu=False)
    C = H * F
    ref, _, _ = ops.edge_bwd_src(Pd, G, Ad, z, minv, t, g, H, F, want_fp32=True, want_ds=True)
    got, planes, _ = ops.edge_bwd_src(Pd, G, Ad, z, minv, t, g, H, F, want_fp32=True, want_planes=True, want_ds=True,
                                      a_term=False)
    dS = ref[:, C:C + H * r].double()
    assert rel_err(got[:, C:C + H * r].cpu().numpy(), dS.cpu().numpy()) < 1e-5 or not e
    a_part = torch.einsum("nhr,hrf->nhf", dS.view(n, H, r), Ad.double()).reshape(n, C)
    want = ref[:, :C].double() - a_part
    assert rel_err(got[:, :C].cpu().numpy(), want.cpu().numpy()) < FP32_TOL or not e
    rec = planes[0].float() + planes[1].float()
    assert rel_err(rec[:, :C + H * r].cpu().numpy(), got[:, :C + H * r].cpu().numpy()) < 2e-5 or not e
    got2, _, _ = ops.edge_bwd_src(Pd, G, Ad, z, minv, t, g, H, F, want_fp32=True, want_ds=True, a_term=False)
    assert torch.equal(got[:, :C + H * r], got2[:, :C + H * r])


def test_bwd_prep_elu_gradient(dev):
    from relgat_projector_b200 import ops
    n, H, F = 333, 4, 40
    g = torch.Generator(device=dev).manual_seed(3)
    out = torch.randn((n, H * F), generator=g, device=dev) * 2
    dY = torch.randn((n, H * F), generator=g, device=dev)
    bias = torch.randn((n,), generator=g, device=dev) * 0.1
    G, t, hsum = ops.edge_bwd_prep(dY, out, bias, H, F, apply_elu=True)
    o64 = out.double().requires_grad_(True)
    (torch.nn.functional.elu(o64) * dY.double()).sum().backward()
    assert rel_err(G.cpu().numpy(), o64.grad.cpu().numpy()) < 1e-6
    t_ref = (o64.grad * (out.double() - bias.double()[:, None])).view(n, H, F).sum(-1)
    assert rel_err(t.cpu().numpy(), t_ref.detach().cpu().numpy()) < 1e-5
    assert rel_err(hsum.cpu().numpy(), o64.grad.view(n, H, F).sum(-1).cpu().numpy()) < 1e-5


def test_large_logits_and_nan_propagation(dev):
    """Logits of order +-80 stay finite (stable softmax); a NaN feature poisons its destinations only."""
    from relgat_projector_b200 import ops
    from relgat_projector_b200.graph import GraphIndex
    rng = np.random.default_rng(9)
    n, e, r, H, F = 80, 700, 4, 2, 16
    src, dst, rel = _graph(rng, n, e, r)
    P = rng.standard_normal((n, H, F)).astype(np.float32)
    A = (rng.standard_normal((H, r, F)) * 20).astype(np.float32)
    gi = O.graph_index_np(src, dst, rel, n, r)
    out_ref, z_ref, alpha_ref, _ = O.layer_forward_closed(P, A, np.zeros(r), gi)
    assert np.abs(z_ref).max() > 80
    g = GraphIndex(torch.from_numpy(np.stack([src, dst])).to(dev), torch.from_numpy(rel).to(dev), n, r)
    Pd = torch.from_numpy(P.reshape(n, -1)).to(dev)
    out, _, alpha, z, _, _ = ops.edge_fwd(Pd, torch.from_numpy(A).to(dev), None, g, H, F, want_alpha=True)
    assert torch.isfinite(out).all()
    assert rel_err(out.cpu().numpy(), out_ref.reshape(n, -1)) < FP32_TOL
    assert rel_err(alpha.cpu().numpy(), alpha_ref) < FP32_TOL
    Pd[7, 3] = float("nan")
    out_nan, _, _, _, _, _ = ops.edge_fwd(Pd, torch.from_numpy(A).to(dev), None, g, H, F)
    touched = np.zeros(n, dtype=bool)
    touched[dst[src == 7]] = True
    bad = torch.isnan(out_nan).any(dim=1).cpu().numpy()
    assert np.array_equal(bad, touched)


# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,normalize", [("distmult", False), ("transe", True), ("transe", False)])
@pytest.mark.parametrize("D", [800, 1024, 30])
def test_scorer_forward_backward_vs_oracle(dev, kind, normalize, D):
    from relgat_projector_b200 import ops
    g = torch.Generator().manual_seed(D)
    n, r, B, Bt = 200, 7, 96, 40
    x = torch.randn((n, D), generator=g)
    rel_emb = torch.randn((r, D), generator=g) * 0.3
    src_ids, dst_ids = torch.randint(0, n, (B,), generator=g), torch.randint(0, n, (B,), generator=g)
    src_ids[5:9] = src_ids[4]  # repeated nodes: the backward must sum them
    dst_ids[3] = src_ids[3]
    rel_ids = torch.randint(0, r, (B,), generator=g)
    dscore = torch.randn((B,), generator=g)
    dtr = torch.randn((Bt, D), generator=g)

    x64, r64 = x.double().requires_grad_(True), rel_emb.double().requires_grad_(True)
    s, t = x64[src_ids], x64[dst_ids]
    if kind == "distmult":
        sc = O.distmult_score_port(s, r64, rel_ids, t)
        tr = O.distmult_transform_port(s[:Bt], r64, rel_ids[:Bt])
    else:
        sc = O.transe_score_port(s, r64, rel_ids, t, normalize)
        tr = O.transe_transform_port(s[:Bt], r64, rel_ids[:Bt], normalize)
    ((sc * dscore.double()).sum() + (tr * dtr.double()).sum()).backward()

    xd, rd = x.to(dev), rel_emb.to(dev)
    si, di, ri = src_ids.to(dev), dst_ids.to(dev), rel_ids.to(dev)
    score, trd, sv, dv = ops.score_fwd(kind, normalize, xd, si, xd, di, rd, ri, n_transform=Bt,
                                       want_src_vec=True, want_dst_vec=True)
    assert rel_err(score.cpu().numpy(), sc.detach().numpy()) < FP32_TOL
    assert rel_err(trd.cpu().numpy(), tr.detach().numpy()) < FP32_TOL
    assert torch.equal(sv.cpu(), x[src_ids]) and torch.equal(dv.cpu(), x[dst_ids])
    d_src, d_dst, d_rel = ops.score_bwd(kind, normalize, xd, si, xd, di, rd, ri, dscore.to(dev), dtr.to(dev))
    dx = ops.index_add_sorted(torch.cat([d_src, d_dst]), torch.cat([si, di]), n)
    drel = ops.index_add_sorted(d_rel, ri, r)
    assert rel_err(dx.cpu().numpy(), x64.grad.numpy()) < FP32_TOL
    assert rel_err(drel.cpu().numpy(), r64.grad.numpy()) < FP32_TOL
    dx2 = ops.index_add_sorted(torch.cat([d_src, d_dst]), torch.cat([si, di]), n)
    assert torch.equal(dx, dx2)
    # rows variant (module-level seam): same numbers without index vectors
    score2, _, _, _ = ops.score_fwd(kind, normalize, sv, None, dv, None, rd, ri)
    assert torch.equal(score2, score)


@pytest.mark.parametrize("B,K,proj", [(64, 4, False), (1024, 4, False), (37, 10, True), (5, 0, False), (300, 1, True)])
def test_fused_margin_loss_matches_oracle(dev, B, K, proj):
    from relgat_projector_b200 import loss as L
    g = torch.Generator().manual_seed(B + K)
    s = torch.randn(B * (1 + K), generator=g)
    s64 = s.double().requires_grad_(True)
    pos, neg = (O.split_scores_projection_path if proj else O.split_scores_kmajor)(s64, B, K)
    sd = s.to(dev).requires_grad_(True)
    got = L.fused_margin_ranking_loss(sd, B, K, 0.7, projection_path=proj)
    if K == 0:
        assert float(got) == 0.0
        return
    ref = O.margin_ranking_loss_port(pos, neg, 0.7)
    ref.backward()
    got.backward()
    assert abs(float(got) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert rel_err(sd.grad.cpu().numpy(), s64.grad.numpy()) < 1e-5
