"""Work tables of the streaming edge kernels (relgat_projector_b200/graph.py::StreamChunks) — host logic, CPU.
Every edge must be processed exactly once, by an ordinary chunk (whole segments) or by one part of a split segment."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from relgat_projector_b200.graph import StreamChunks


def _check(deg, chunk_edges, chunk_nodes, long_segment, part_edges):
    deg = np.asarray(deg, dtype=np.int64)
    ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]).astype(np.int32))
    ck = StreamChunks(ptr, chunk_edges=chunk_edges, chunk_nodes=chunk_nodes, long_segment=long_segment,
                      part_edges=part_edges)
    n, e = len(deg), int(deg.sum())
    chunks, parts = ck.chunks.numpy(), ck.parts.numpy()
    assert ck.n_chunks == len(chunks) and ck.n_parts == len(parts) and ck.n_long == len(ck.long_node)
    seg_seen = np.zeros(n, dtype=np.int64)
    edge_seen = np.zeros(e, dtype=np.int64)
    p = ptr.numpy().astype(np.int64)
    seen_ordinary = False
    for lo, nn, part, _ in chunks:
        assert 1 <= nn <= chunk_nodes
        if part < 0:
            seen_ordinary = True
            seg_seen[lo:lo + nn] += 1
            edge_seen[p[lo]:p[lo + nn]] += 1
            assert (deg[lo:lo + nn] <= long_segment).all()
        else:
            assert not seen_ordinary, "parts of split segments are scheduled first"
            assert nn == 1 and deg[lo] > long_segment
            a, b = parts[part]
            assert p[lo] <= a < b <= p[lo + 1] and b - a <= part_edges
            edge_seen[a:b] += 1
    long_nodes = ck.long_node.numpy()
    assert (np.sort(long_nodes) == np.nonzero(deg > long_segment)[0]).all()
    seg_seen[long_nodes] += 1
    assert (seg_seen == 1).all(), "every segment (also empty ones) belongs to exactly one chunk or is split"
    assert (edge_seen == 1).all()
    lp = ck.long_part_ptr.numpy()
    assert lp[0] == 0 and lp[-1] == ck.n_parts and len(lp) == ck.n_long + 1
    for i, node in enumerate(long_nodes):  # parts of one segment are consecutive, ordered, and tile it exactly
        seg = parts[lp[i]:lp[i + 1]]
        assert seg[0][0] == p[node] and seg[-1][1] == p[node + 1]
        assert (seg[1:, 0] == seg[:-1, 1]).all()


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 40), min_size=0, max_size=120), st.integers(1, 16), st.integers(1, 8),
       st.integers(4, 24), st.integers(1, 8))
def test_stream_chunks_cover_every_edge_once(deg, chunk_edges, chunk_nodes, long_segment, part_edges):
    _check(deg, chunk_edges, chunk_nodes, long_segment, part_edges)


@pytest.mark.parametrize("deg", [[], [0], [0, 0, 0], [5000], [0, 600, 0, 3, 513, 512], [1] * 1000])
def test_stream_chunks_edge_cases_default_sizes(deg):
    _check(deg, 32, 64, 512, 256)
