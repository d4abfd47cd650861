"""CPU replay of the peer-table partition (relgat_projector_b200/peer.py::PeerIndexPlan) with the numpy oracle:
every rank's forward (by destination owner) and backward (by source owner) is recomputed from the plan's index
arrays and from emulated "mapped ranges", and must reproduce the whole-graph closed form."""
import numpy as np
import pytest
import torch

from oracle import relgat_oracle as O
from relgat_projector_b200 import peer as RP


def _graph(seed, n, e, r):
    rng = np.random.default_rng(seed)
    dst = (rng.zipf(1.6, size=e) % n).astype(np.int64)  # skewed destinations: uneven ranges and halos
    src = rng.integers(0, n, size=e).astype(np.int64)
    src = np.where(src == dst, (src + 1) % n, src)
    rel = rng.integers(0, r, size=e).astype(np.int64)
    return src, dst, rel


def _mapped_view(tabs, rank, world):
    """What rank `rank` sees: all ranks' tables back to back, its own first (peer_table.cu: slot s = rank + s)."""
    return np.concatenate([tabs[(rank + s) % world] for s in range(world)], 0)


@pytest.mark.parametrize("world,blocks", [(2, 1), (3, 2), (4, 3)])
def test_plan_replays_whole_graph_layer(world, blocks):
    n, e, r, h, f = 83, 900, 5, 3, 4
    src, dst, rel = _graph(world, n, e, r)
    rng = np.random.default_rng(7)
    P, G = rng.standard_normal((n, h, f)), rng.standard_normal((n, h, f))
    A, beta = rng.standard_normal((h, r, f)) * 0.5, rng.standard_normal(r) * 0.1
    g = O.graph_index_np(src, dst, rel, n, r)
    out, z, alpha, bias = O.layer_forward_closed(P, A, beta, g)
    dP, dA, dbeta, dz = O.layer_backward_closed(G, P, A, g, z, alpha)
    t = np.einsum("nhf,nhf->nh", G, out - bias[:, None, None])  # the delta identity the kernels use
    ei = torch.from_numpy(np.stack([src, dst]))
    et = torch.from_numpy(rel)
    plans = [RP.PeerIndexPlan(ei, et, n, rk, world, [(o - rk) % world for o in range(world)],
                              lambda m, b: max(m, 1) + 3, [16], [16], blocks=blocks) for rk in range(world)]
    stride, slots = plans[0].stride_rows, plans[0].stride_slots
    assert all(p.stride_rows == stride and p.stride_slots == slots and p.bounds == plans[0].bounds for p in plans)
    assert sum(p.E_fwd for p in plans) == sum(p.E_bwd for p in plans) == e

    def table(rows_of_rank, width_shape, rows=stride):
        tabs = []
        for p in plans:
            tb = np.zeros((rows,) + width_shape)
            vals = rows_of_rank(p)
            tb[: len(vals)] = vals
            tabs.append(tb)
        return tabs

    P_tabs = table(lambda p: P[p.lo:p.hi], (h, f))
    # ---------------- forward, by destination owner ----------------
    z_tabs, a_tabs = [], []
    for rk, p in enumerate(plans):
        nl = p.n_local
        view = _mapped_view(P_tabs, rk, world)
        P_ext = np.concatenate([P[p.lo:p.hi], view[p.pull_f.numpy()]], 0)
        fs, fd, fr = (x.numpy() for x in p.fwd_edges)
        assert fs.max(initial=0) < nl + p.n_halo_f and fd.max(initial=0) < max(nl, 1)
        gl = O.graph_index_np(fs, fd, fr, nl + p.n_halo_f, r)
        out_l, z_l, alpha_l, _ = O.layer_forward_closed(P_ext, A, beta, gl)
        np.testing.assert_allclose(out_l[:nl], out[p.lo:p.hi], rtol=0, atol=1e-12)
        lo_e, hi_e = g["rowptr"][p.lo], g["rowptr"][p.hi]
        np.testing.assert_array_equal(z_l, z[lo_e:hi_e])  # local CSR order == the global one restricted to the range
        for tabs, vals in ((z_tabs, z_l), (a_tabs, alpha_l)):
            tb = np.zeros((slots, h))
            tb[: len(vals)] = vals
            tabs.append(tb)
        # pulled rows of block c come from block c of their owner; the first rows of a block rotate over the owners
        off = 0
        for c, cnt in enumerate(p.blk_f):
            ids = p.pull_f.numpy()[off:off + cnt]
            slot, row = ids // stride, ids % stride
            owner = (rk + slot) % world
            size = np.array([plans[o].n_local for o in owner])
            assert ((row * blocks) // np.maximum(size, 1) == c).all()
            distinct = len(set(owner.tolist()))
            assert len(set(owner[:distinct].tolist())) == distinct or cnt < distinct
            off += cnt
    # ---------------- backward, by source owner ----------------
    G_tabs = table(lambda p: G[p.lo:p.hi], (h, f))
    t_tabs = table(lambda p: t[p.lo:p.hi], (h,))
    dA_sum, dbeta_sum = np.zeros_like(dA), np.zeros_like(dbeta)
    for rk, p in enumerate(plans):
        nl = p.n_local
        pull_b = p.pull_b.numpy()
        G_ext = np.concatenate([G[p.lo:p.hi], _mapped_view(G_tabs, rk, world)[pull_b]], 0)
        t_ext = np.concatenate([t[p.lo:p.hi], _mapped_view(t_tabs, rk, world)[pull_b]], 0)
        bs, bd, br = (x.numpy() for x in p.bwd_edges)
        zr = p.z_row_bwd.numpy()
        z_e = _mapped_view(z_tabs, rk, world)[zr]
        a_e = _mapped_view(a_tabs, rk, world)[zr]
        P_own = P[p.lo:p.hi]
        dalpha = np.einsum("ehf,ehf->eh", G_ext[bd], P_own[bs])
        dz_e = a_e * (dalpha - t_ext[bd]) * np.where(z_e > 0, 1.0, O.LEAKY_SLOPE)
        dP_l = np.zeros((nl, h, f))
        np.add.at(dP_l, bs, a_e[:, :, None] * G_ext[bd] + dz_e[:, :, None] * np.transpose(A[:, br, :], (1, 0, 2)))
        np.testing.assert_allclose(dP_l, dP[p.lo:p.hi], rtol=0, atol=1e-10)  # complete on its owner: no cross-rank sum
        for hh in range(h):
            np.add.at(dA_sum[hh], br, dz_e[:, hh, None] * P_own[bs][:, hh, :])
        np.add.at(dbeta_sum, br, G_ext[bd].sum(axis=(1, 2)))
    np.testing.assert_allclose(dA_sum, dA, rtol=0, atol=1e-10)
    np.testing.assert_allclose(dbeta_sum, dbeta, rtol=0, atol=1e-10)


def test_plan_world1_and_empty_rank():
    """world 1: no halo, identity numbering; a rank whose range holds no destination of any edge still gets a plan."""
    src, dst, rel = _graph(3, 40, 200, 4)
    ei, et = torch.from_numpy(np.stack([src, dst])), torch.from_numpy(rel)
    p = RP.PeerIndexPlan(ei, et, 40, 0, 1, [0], lambda m, b: max(m, 1), [16], [16], blocks=2)
    assert p.n_halo_f == p.n_halo_b == 0 and p.E_fwd == p.E_bwd == 200
    assert torch.equal(p.fwd_edges[0], ei[0]) and torch.equal(p.bwd_edges[1], ei[1])
    dst2 = np.zeros_like(dst)  # every edge points at node 0: the upper ranges own no in-edge
    ei2 = torch.from_numpy(np.stack([np.where(src == 0, 1, src), dst2]))
    plans = [RP.PeerIndexPlan(ei2, et, 40, rk, 2, [(o - rk) % 2 for o in range(2)], lambda m, b: max(m, 1), [16], [16],
                              balance="nodes") for rk in range(2)]
    assert plans[1].E_fwd == 0 and plans[0].E_fwd == 200 and plans[0].E_bwd + plans[1].E_bwd == 200


@pytest.mark.parametrize("world", [2, 4])
def test_sparse_halo_targets_equal_the_dense_pull(world):
    """Batch-sparse last layer: scattering only the batch rows that sit in a rank's backward halo into a zeroed pulled
    block gives the same [own | pulled] table as pulling the whole (almost-all-zero) halo."""
    n, e, r, w = 90, 1000, 4, 5
    src, dst, rel = _graph(11 + world, n, e, r)
    ei, et = torch.from_numpy(np.stack([src, dst])), torch.from_numpy(rel)
    plans = [RP.PeerIndexPlan(ei, et, n, rk, world, [(o - rk) % world for o in range(world)],
                              lambda m, b: max(m, 1), [16], [16], blocks=2) for rk in range(world)]
    stride = plans[0].stride_rows
    rng = np.random.default_rng(5)
    ids = rng.integers(0, n, size=40)  # with repeats
    G = np.zeros((n, w))
    G[np.unique(ids)] = rng.standard_normal((len(np.unique(ids)), w))
    tabs = []
    for p in plans:
        tb = np.zeros((stride, w))
        tb[: p.n_local] = G[p.lo:p.hi]
        tabs.append(tb)
    for rk, p in enumerate(plans):
        view = _mapped_view(tabs, rk, world)
        dense = np.concatenate([G[p.lo:p.hi], view[p.pull_b.numpy()]], 0)
        src_rows, dst_rows = (x.numpy() for x in p.sparse_halo_targets(torch.from_numpy(ids)))
        assert len(src_rows) == len(ids) and dst_rows.max() <= p.n_local + p.n_halo_b < stride
        local = np.zeros((stride, w))
        local[: p.n_local] = G[p.lo:p.hi]
        local[dst_rows] = view[src_rows]
        np.testing.assert_array_equal(local[: p.n_local + p.n_halo_b], dense)
