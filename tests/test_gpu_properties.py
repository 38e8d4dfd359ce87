"""Property tests (hypothesis) of the bit-exact index work on the device — COO -> CSR, CSR transpose, index-block
transpose, dense mask -> CSR — against numpy's stable sorts, on ragged / empty / duplicated inputs."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from graphneuralnetwork_b200.graph import CSRGraph, index_block_transpose

pytestmark = pytest.mark.gpu
DEV = "cuda"
SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@st.composite
def coo(draw):
    n_rows = draw(st.integers(1, 70))
    n_cols = draw(st.integers(1, 70))
    nnz = draw(st.integers(0, 400))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    hot = draw(st.booleans())  # a hub row / many duplicates
    row = rng.integers(0, n_rows, nnz)
    if hot and nnz:
        row[rng.random(nnz) < 0.5] = rng.integers(0, n_rows)
    col = rng.integers(0, n_cols, nnz)
    val = rng.standard_normal(nnz).astype(np.float32)
    return n_rows, n_cols, row.astype(np.int64), col.astype(np.int64), val


@settings(**SETTINGS)
@given(coo())
def test_coo_to_csr_is_the_stable_sort_by_row(lib, g):
    n_rows, n_cols, row, col, val = g
    csr, perm = CSRGraph.from_coo(cuda(row), cuda(col), cuda(val), n_rows, n_cols, return_perm=True)
    order = np.argsort(row, kind="stable")
    assert np.array_equal(perm.cpu().numpy(), order)
    assert np.array_equal(csr.col.cpu().numpy(), col[order].astype(np.int32))
    assert np.array_equal(csr.val.cpu().numpy().view(np.uint32), val[order].view(np.uint32))
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(row, minlength=n_rows))])
    assert np.array_equal(csr.rowptr.cpu().numpy(), rowptr)


@settings(**SETTINGS)
@given(coo())
def test_csr_transpose_is_stable_and_an_involution(lib, g):
    n_rows, n_cols, row, col, val = g
    csr = CSRGraph.from_coo(cuda(row), cuda(col), cuda(val), n_rows, n_cols)
    t = csr.transpose()
    r_sorted = np.sort(row, kind="stable")
    order = np.argsort(row, kind="stable")
    c_sorted, v_sorted = col[order], val[order]
    order_t = np.argsort(c_sorted, kind="stable")  # transposed rows list their sources in ascending original order
    assert np.array_equal(t.col.cpu().numpy(), r_sorted[order_t].astype(np.int32))
    assert np.array_equal(t.val.cpu().numpy().view(np.uint32), v_sorted[order_t].view(np.uint32))
    assert np.array_equal(csr.perm_t.cpu().numpy(), order_t)
    tt = CSRGraph(t.rowptr, t.col, t.val, n_cols, n_rows).transpose()
    # transposing twice restores the pattern (edge order within a row becomes ascending-column, stable)
    back = sorted(zip(tt.edge_rows().cpu().tolist(), tt.col.cpu().tolist()))
    assert back == sorted(zip(row.tolist(), col.tolist()))


@settings(**SETTINGS)
@given(st.integers(1, 60), st.integers(1, 9), st.integers(1, 80), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_index_block_transpose_lists_positions_in_order(lib, n_src, fanout, n_table, seed, use_i32):
    rng = np.random.default_rng(seed)
    idx = rng.integers(-1, n_table, n_src * fanout)  # -1 = padding, skipped
    t = cuda(idx.astype(np.int32 if use_i32 else np.int64))
    rowptr_t, pos_t = index_block_transpose(t, n_table, cache=False)
    rp = rowptr_t.cpu().numpy()
    pos = pos_t.cpu().numpy()
    for r in range(n_table):
        assert np.array_equal(pos[rp[r]:rp[r + 1]], np.nonzero(idx == r)[0])
    assert rp[-1] == int((idx >= 0).sum())


@settings(**SETTINGS)
@given(st.integers(1, 50), st.integers(1, 50), st.floats(0.0, 1.0), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_dense_mask_to_csr_matches_nonzero_order(lib, n_rows, n_cols, density, seed, f64):
    rng = np.random.default_rng(seed)
    adj = (rng.random((n_rows, n_cols)) < density) * rng.random((n_rows, n_cols))
    adj[rng.random((n_rows, n_cols)) < 0.05] = -1.0  # negative entries are NOT edges (`adj > 0`)
    adj = adj.astype(np.float64 if f64 else np.float32)
    csr = CSRGraph.from_dense_mask(cuda(adj))
    r, c = np.nonzero(adj > 0)
    assert np.array_equal(csr.col.cpu().numpy(), c.astype(np.int32))
    assert np.array_equal(csr.rowptr.cpu().numpy(), np.concatenate([[0], np.cumsum(np.bincount(r, minlength=n_rows))]))


def test_index_transpose_cache_is_identity_guarded(lib):
    idx = torch.randint(0, 50, (200,), device=DEV)
    a = index_block_transpose(idx, 50)
    assert index_block_transpose(idx, 50)[0] is a[0]          # same tensor: served from the cache
    idx[0] = (idx[0] + 1) % 50                                # in-place edit bumps the version
    b = index_block_transpose(idx, 50)
    assert b[0] is not a[0]
    ptr = idx.data_ptr()
    del idx
    fresh = torch.randint(0, 50, (200,), device=DEV)          # usually lands on the freed block
    c = index_block_transpose(fresh, 50)
    want = torch.bincount(fresh, minlength=50).cumsum(0)
    assert torch.equal(c[0][1:], want) and (fresh.data_ptr() != ptr or c[0] is not b[0])
