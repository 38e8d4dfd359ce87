"""Seeded synthetic inputs of the BASELINE.json shapes.  The reference ships
no datasets and there is no network, so every config is synthetic.  Host-side generators
are numpy; the papers100M/Reddit-scale graphs are generated on the device through the C ABI.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .graph import CSRGraph, _p, _stream_ptr

CORA = dict(n=2708, undirected_pairs=5278, feats=1433, classes=7, hidden=16)
REDDIT = dict(n=232_965, edges=114_615_892, feats=602, classes=41, batch=1024, fanout=(25, 10), hidden=(128, 41))
ACM = dict(n=3025, feats=1870, classes=3, metapath_nnz=(29_000, 2_210_000, 300_000), heads=8, hidden=8)
PAPERS100M = dict(n=111_059_956, edges=1_615_685_872, feats=128)


def cora_like_edges(n=CORA["n"], pairs=CORA["undirected_pairs"], seed=0) -> np.ndarray:
    """`pairs` distinct undirected pairs (i<j), uniform, returned as a directed [pairs,2] int32
    list — the shape of a .cites file before GCN/data_utils.py:35 symmetrises it."""
    rng = np.random.default_rng(seed)
    got = set()
    out = np.empty((pairs, 2), dtype=np.int32)
    k = 0
    while k < pairs:
        i, j = rng.integers(0, n, size=2)
        if i == j:
            continue
        a, b = (int(i), int(j)) if i < j else (int(j), int(i))
        if (a, b) in got:
            continue
        got.add((a, b))
        out[k] = (a, b)
        k += 1
    return out


def row_normalised_features(n, f, seed=0) -> np.ndarray:
    """rand features, row-normalised like GCN/data_utils.py:39-51 leaves them."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, f), dtype=np.float32)
    return x / x.sum(1, keepdims=True)


def symmetric_mask(n, target_nnz, seed=0, dtype=np.float64) -> np.ndarray:
    """Symmetric 0/1 adjacency with self-loops and about `target_nnz` non-zeros — the dense
    float64 metapath adjacency of HAN/utils/data_utils.py:85-89."""
    rng = np.random.default_rng(seed)
    pairs = max((target_nnz - n) // 2, 0)
    adj = np.zeros((n, n), dtype=dtype)
    i = rng.integers(0, n, size=pairs)
    j = rng.integers(0, n, size=pairs)
    adj[i, j] = 1
    adj[j, i] = 1
    adj[np.arange(n), np.arange(n)] = 1
    return adj


def adjacency_lists(n, avg_degree, seed=0, min_degree=1):
    """dict node -> set(neighbours) (the `adj_lists` of GraphSAGE_Pytorch/data_utils.py:31-40),
    undirected, skewed degrees."""
    rng = np.random.default_rng(seed)
    m = int(n * avg_degree / 2)
    w = 1.0 / np.arange(1, n + 1) ** 0.5
    w /= w.sum()
    a = rng.choice(n, size=m, p=w)
    b = rng.integers(0, n, size=m)
    adj = {i: set() for i in range(n)}
    for x, y in zip(a.tolist(), b.tolist()):
        if x != y:
            adj[x].add(y)
            adj[y].add(x)
    for i in range(n):
        while len(adj[i]) < min_degree:
            j = int(rng.integers(0, n))
            if j != i:
                adj[i].add(j)
                adj[j].add(i)
    return adj


def uniform_blocks(n_nodes, batch, fanouts, seed=0, dtype=torch.int64, device="cpu", generator=None):
    """Throughput-run index blocks: uniform ids of lengths B, B·f1, B·f1·f2."""
    g = generator or torch.Generator(device="cpu").manual_seed(seed)
    sizes = [batch]
    for f in fanouts:
        sizes.append(sizes[-1] * f)
    return [torch.randint(0, n_nodes, (s,), generator=g, dtype=torch.int64).to(dtype).to(device) for s in sizes]


def powerlaw_csr(n_rows: int, mean_degree: float, *, n_cols: int | None = None, row_offset: int = 0,
                 exponent: float = 2.5, skew: float = 3.0, max_degree: int = 1 << 20, seed: int = 0,
                 device="cuda", with_values: bool = True, deg_all: torch.Tensor | None = None,
                 p_local: float = 0.0, window: int = 0, scatter_hubs: bool = False) -> CSRGraph:
    """Power-law CSR generated on the device (gnn_synth_*): Pareto degrees with the requested
    mean, neighbour ids skewed to low ids (hubs), one self-loop per row, GCN-normalised
    values d_i^-1/2·d_j^-1/2 from the row degrees.  `row_offset`/`n_cols` generate one row
    block of a larger graph (each rank of the partitioned run builds only its own rows).
    `p_local` > 0 draws that share of the edges within +-`window` of the row id instead — the
    structure a locality-preserving (METIS-like) node ordering gives a 1-D partition.
    `scatter_hubs` maps the popular (hub) ids through a fixed bijection of the id range, so
    hubs are spread over all row blocks as in a real graph instead of sitting at the low ids."""
    if scatter_hubs:
        window = -max(int(window), 1)
    lib = _lib.load()
    n_cols = n_rows if n_cols is None else n_cols
    dev = torch.device(device)
    deg = torch.empty(n_rows, dtype=torch.int64, device=dev)
    _lib.check(lib.gnn_synth_powerlaw_degrees(n_rows, row_offset, float(mean_degree), float(exponent), int(max_degree),
                                              seed, _p(deg), _stream_ptr()), "gnn_synth_powerlaw_degrees")
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=rowptr[1:])
    nnz = int(rowptr[-1].item())
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    _lib.check(lib.gnn_synth_powerlaw_fill(n_rows, row_offset, n_cols, _p(rowptr), float(skew), float(p_local),
                                           int(window), seed, _p(col), _stream_ptr()), "gnn_synth_powerlaw_fill")
    val = None
    if with_values:
        if deg_all is None:
            if n_cols != n_rows or row_offset != 0:
                deg_all = torch.empty(n_cols, dtype=torch.int64, device=dev)
                _lib.check(lib.gnn_synth_powerlaw_degrees(n_cols, 0, float(mean_degree), float(exponent),
                                                          int(max_degree), seed, _p(deg_all), _stream_ptr()),
                           "gnn_synth_powerlaw_degrees")
            else:
                deg_all = deg
        val = torch.empty(nnz, dtype=torch.float32, device=dev)
        _lib.check(lib.gnn_synth_gcn_values(n_rows, row_offset, _p(rowptr), _p(col), _p(deg_all), _p(val),
                                            _stream_ptr()), "gnn_synth_gcn_values")
    del deg
    return CSRGraph(rowptr, col, val, n_rows, n_cols)


def hashed_features(rows: torch.Tensor, F: int, dtype=torch.float32, salt: int = 0) -> torch.Tensor:
    """Feature rows that are a pure function of the GLOBAL row id: x[i, f] = g(i * F + f) in (-1, 1), exact
    integer arithmetic (two rounds of a Lehmer generator), so any rank — and a float64 checker — can
    recompute any row of X without communication (bench.py's partitioned-SpMM correctness check)."""
    M = 2147483647
    z = (rows.to(torch.int64).view(-1, 1) * F + torch.arange(F, dtype=torch.int64, device=rows.device) + salt) % M
    z = (z * 48271 + 11) % M
    z = (z * 69621 + 7) % M
    return (z.to(torch.float64) * (2.0 / M) - 1.0).to(dtype)


def hashed_feature_block(lo: int, hi: int, F: int, device, dtype=torch.float32, salt: int = 0,
                         rows_per_pass: int = 1 << 21, out: torch.Tensor | None = None) -> torch.Tensor:
    """hashed_features for the contiguous global rows [lo, hi), generated in passes (bounded scratch)."""
    X = out if out is not None else torch.empty((hi - lo, F), dtype=dtype, device=device)
    for a in range(lo, hi, rows_per_pass):
        b = min(a + rows_per_pass, hi)
        X[a - lo:b - lo] = hashed_features(torch.arange(a, b, dtype=torch.int64, device=device), F, dtype, salt)
    return X


def spmm_check_rows(csr: CSRGraph, n_samples: int, seed: int) -> torch.Tensor:
    """Row sample for the float64 spot check: uniform rows plus the longest ones (chunked path)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    rows = torch.randint(0, csr.n_rows, (n_samples,), generator=g).to(csr.rowptr.device)
    deg = csr.rowptr[1:] - csr.rowptr[:-1]
    hubs = torch.topk(deg, min(8, csr.n_rows)).indices
    return torch.unique(torch.cat([rows, hubs]))


def spmm_sampled_reference(csr: CSRGraph, rows: torch.Tensor, F: int, dtype=torch.float32, salt: int = 0,
                           with_abs: bool = False):
    """float64 rows `rows` of Â·X where X = hashed_features of the GLOBAL column ids held in `csr.col`
    (recomputed here, no communication) — torch float64 arithmetic in plain CSR order."""
    from .partition import select_rows
    rp, cc, vv = select_rows(csr.rowptr, csr.col, csr.val, rows)
    src = hashed_features(cc.to(torch.int64), F, dtype, salt).double()
    if vv is not None:
        src *= vv.double().unsqueeze(1)
    owner = torch.repeat_interleave(torch.arange(rows.numel(), device=rows.device), rp[1:] - rp[:-1])
    ref = torch.zeros((rows.numel(), F), dtype=torch.float64, device=rows.device)
    ref.index_add_(0, owner, src)
    if with_abs:
        # Σ_j |a_ij x_jf|: the scale an fp32 sum's rounding error is relative to (a hub row adds 10^5 terms)
        mag = torch.zeros_like(ref)
        mag.index_add_(0, owner, src.abs())
        return ref, mag
    return ref
