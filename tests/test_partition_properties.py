"""Hypothesis property tests (CPU) of the pure host functions under the partitioned SpMM plan
(graphneuralnetwork_b200/partition.py; SURVEY.md §4 "property tests", §8e): nnz-balanced row blocks, the
local / remote column split and the compact row-subset CSR — on ragged graphs with empty rows, hub rows,
duplicate edges, empty blocks and more ranks than rows."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from graphneuralnetwork_b200.partition import balanced_bounds, choose_two_pass_chunks, select_rows, split_columns


@st.composite
def csr_graphs(draw, max_rows=40, max_deg=12):
    n = draw(st.integers(1, max_rows))
    deg = draw(st.lists(st.integers(0, max_deg), min_size=n, max_size=n))
    if draw(st.booleans()):
        deg[draw(st.integers(0, n - 1))] += draw(st.integers(0, 60))  # a hub row
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    nnz = int(rowptr[-1])
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    col = rng.integers(0, n, nnz).astype(np.int64)               # duplicates allowed, any order within a row
    val = rng.standard_normal(nnz).astype(np.float32)
    return n, rowptr, col, val


@settings(max_examples=150, deadline=None)
@given(csr_graphs(), st.integers(1, 9))
def test_balanced_bounds_partition_the_rows(g, world):
    n, rowptr, _, _ = g
    b = balanced_bounds(torch.from_numpy(rowptr), world)
    assert len(b) == world + 1 and b[0] == 0 and b[-1] == n
    assert all(x <= y for x, y in zip(b, b[1:]))                 # monotone: empty blocks allowed, no overlap
    total, max_deg = int(rowptr[-1]), int(np.diff(rowptr).max())
    for p in range(world):                                       # every block within one row of its nnz share
        nnz_p = int(rowptr[b[p + 1]] - rowptr[b[p]])
        assert nnz_p <= total // world + max_deg + 1


@settings(max_examples=150, deadline=None)
@given(csr_graphs(), st.integers(1, 5), st.data())
def test_split_columns_is_a_stable_partition_of_every_row(g, world, data):
    n, rowptr, col, val = g
    b = balanced_bounds(torch.from_numpy(rowptr), world)
    rank = data.draw(st.integers(0, world - 1))
    lo, hi = b[rank], b[rank + 1]
    rp = torch.from_numpy(rowptr[lo:hi + 1] - rowptr[lo])
    c = torch.from_numpy(col[rowptr[lo]:rowptr[hi]])
    v = torch.from_numpy(val[rowptr[lo]:rowptr[hi]])
    rl, cl, vl, rr, cr, vr = split_columns(rp, c, v, lo, hi)
    assert rl[-1] + rr[-1] == rp[-1] and cl.dtype == torch.int32
    for i in range(hi - lo):
        row_c, row_v = c[rp[i]:rp[i + 1]].numpy(), v[rp[i]:rp[i + 1]].numpy()
        own = (row_c >= lo) & (row_c < hi)
        assert np.array_equal(cl[rl[i]:rl[i + 1]].numpy(), (row_c[own] - lo).astype(np.int32))   # re-based, in order
        assert np.array_equal(vl[rl[i]:rl[i + 1]].numpy(), row_v[own])
        assert np.array_equal(cr[rr[i]:rr[i + 1]].numpy(), row_c[~own])                          # global ids, in order
        assert np.array_equal(vr[rr[i]:rr[i + 1]].numpy(), row_v[~own])


@settings(max_examples=150, deadline=None)
@given(csr_graphs(), st.data())
def test_select_rows_equals_row_slicing(g, data):
    n, rowptr, col, val = g
    rows = sorted(data.draw(st.sets(st.integers(0, n - 1), max_size=n)))
    with_val = data.draw(st.booleans())
    rp, c, v = select_rows(torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(val) if with_val else None,
                           torch.tensor(rows, dtype=torch.int64))
    assert rp.numel() == len(rows) + 1 and int(rp[0]) == 0 and (v is None) == (not with_val)
    for k, r in enumerate(rows):
        assert np.array_equal(c[rp[k]:rp[k + 1]].numpy(), col[rowptr[r]:rowptr[r + 1]])
        if with_val:
            assert np.array_equal(v[rp[k]:rp[k + 1]].numpy(), val[rowptr[r]:rowptr[r + 1]])


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 8), st.integers(1, 602), st.sampled_from([2, 4]), st.integers(0, 10 ** 8), st.integers(0, 10 ** 7),
       st.integers(0, 2 ** 31 - 1), st.sampled_from([2, 4, 8]))
def test_two_pass_model_is_total_and_bounded(K, F, elem, nnz_p1, rows_int, seed, world):
    """The byte model returns a c0 in [0, K] and a finite modelled time for every c0, for any statistics
    (zero-sized waves, empty chunks); more exchange bytes never make the modelled step shorter."""
    rng = np.random.default_rng(seed)
    m_rows = rng.integers(0, 10 ** 6, K).tolist()
    m_loc = rng.integers(0, 10 ** 7, K).tolist()
    m_rem = rng.integers(0, 10 ** 7, K).tolist()
    waves = (rng.integers(0, 10 ** 9, K) * rng.integers(0, 2, K)).astype(float).tolist()
    c0, rep = choose_two_pass_chunks(K, F, elem, nnz_p1, rows_int, m_rows, m_loc, m_rem, waves, world)
    assert 0 <= c0 <= K and all(np.isfinite(rep[f"c0={c}"]) and rep[f"c0={c}"] >= 0 for c in range(K + 1))
    assert rep[f"c0={c0}"] <= min(rep[f"c0={c}"] for c in range(K + 1)) + 1e-6
    assert rep["exchange_ms"] >= 0 and rep["compute_ms"] >= 0
    _, rep2 = choose_two_pass_chunks(K, F, elem, nnz_p1, rows_int, m_rows, m_loc, m_rem, [2 * w for w in waves], world)
    assert min(rep2[f"c0={c}"] for c in range(K + 1)) >= min(rep[f"c0={c}"] for c in range(K + 1)) - 1e-6
