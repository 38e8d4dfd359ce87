"""Single-GPU parity of the entry points the partitioned SpMM is built from, through the
C ABI: the row-subset / two-table SpMM (gnn_spmm_csr_ex_*), both halo movers (gnn_halo_push, with local
buffers standing in for the peers' — the multi-rank path itself is covered by tools/spmm_dist.py on 2-8
GPUs and by the gloo plan tests on CPU) and the arrival flags.  Bit-exact where only bytes move."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from conftest import rel_err
from graphneuralnetwork_b200 import _lib, functional as Fn
from graphneuralnetwork_b200.graph import CSRGraph
from graphneuralnetwork_b200.partition import HaloPlan, PartitionedSpmm, build_halo_plan, select_rows

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(DEV) if dtype is None else t.to(DEV, dtype)


def _graph(n_rows, n_cols, avg_deg, seed, long_row=None):
    rng = np.random.default_rng(seed)
    deg = rng.poisson(avg_deg, size=n_rows)
    deg[::9] = 0
    if long_row is not None:
        deg[long_row[0]] = long_row[1]
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = rng.integers(0, n_cols, size=rowptr[-1]).astype(np.int32)
    val = rng.standard_normal(rowptr[-1]).astype(np.float32)
    return rowptr, col, val


@pytest.mark.parametrize("F", [16, 128, 130, 602])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_spmm_ex_row_subset_two_tables_and_accumulate_prefix(lib, F, dtype):
    """One launch: compact rows [0, k) add their remote part into Y, the others are written once from
    [X ; X2] — against a float64 scipy product."""
    n_rows, n_loc, n_halo = 3000, 2500, 700
    rowptr, col, val = _graph(n_rows, n_loc + n_halo, 9, seed=F, long_row=(1234, 5000))
    rng = np.random.default_rng(1)
    X = rng.standard_normal((n_loc, F)).astype(np.float32)
    H = rng.standard_normal((n_halo, F)).astype(np.float32)
    Y0 = rng.standard_normal((n_rows, F)).astype(np.float32)
    if dtype == torch.bfloat16:  # the checker sees the same rounded inputs
        X, H, Y0 = [torch.from_numpy(a).bfloat16().float().numpy() for a in (X, H, Y0)]
    rows = np.sort(rng.choice(n_rows, size=1700, replace=False))
    rows = np.union1d(rows, [1234])  # the long row takes the chunked path
    k = len(rows) // 3
    rp, cc, vv = select_rows(torch.from_numpy(rowptr), torch.from_numpy(col), torch.from_numpy(val), torch.from_numpy(rows))
    g = CSRGraph(rp.to(DEV), cc.to(DEV), vv.to(DEV), len(rows), n_loc + n_halo)
    Y = cuda(Y0, dtype)
    Xd, Hd = Fn._padded_empty(n_loc, F, dtype, DEV), Fn._padded_empty(n_halo, F, dtype, DEV)
    Xd.copy_(cuda(X, dtype))
    Hd.copy_(cuda(H, dtype))
    Fn.spmm_ex(g, Xd, Y, row_map=cuda(rows.astype(np.int32)), accumulate_prefix=k, X2=Hd, split=n_loc)
    A = sp.csr_matrix((val.astype(np.float64), col, rowptr), shape=(n_rows, n_loc + n_halo))
    full = A @ np.vstack([X, H]).astype(np.float64)
    ref = Y0.astype(np.float64).copy()
    ref[rows[:k]] += full[rows[:k]]
    ref[rows[k:]] = full[rows[k:]]
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel_err(Y.float().cpu().numpy(), ref) < tol
    untouched = np.setdiff1d(np.arange(n_rows), rows)
    assert np.array_equal(Y.float().cpu().numpy()[untouched], Y0[untouched])  # rows outside the subset are not written


def test_spmm_ex_plain_equals_planned_bitwise(lib):
    """No row_map / no second table: the extended entry point runs the same kernels as the planned one."""
    rowptr, col, val = _graph(4000, 4000, 14, seed=3, long_row=(77, 6000))
    g = CSRGraph(cuda(rowptr), cuda(col), cuda(val), 4000, 4000)
    X = torch.randn(4000, 128, device=DEV)
    a = Fn.spmm_raw(g, X)
    b = Fn.spmm_ex(g, X, torch.empty_like(a), exclusion_smem=28 * 1024)
    assert torch.equal(a, b)


def test_spmm_ex_rejects_bad_arguments(lib):
    rowptr, col, val = _graph(100, 100, 5, seed=4)
    g = CSRGraph(cuda(rowptr), cuda(col), cuda(val), 100, 100)
    X = torch.randn(100, 16, device=DEV)
    with pytest.raises(_lib.GnnError):
        Fn.spmm_ex(g, X, torch.empty(100, 16, device=DEV), row_map=torch.zeros(5, dtype=torch.int32, device=DEV))
    with pytest.raises(_lib.GnnError):
        Fn.spmm_ex(g, X, torch.empty(100, 16, device=DEV), exclusion_smem=64 * 1024)
    o = _lib.SpmmOpts()  # struct_size left 0: version skew must be refused, not read past
    st = lib.gnn_spmm_csr_ex_f32(g.rowptr.data_ptr(), g.col.data_ptr(), g.val.data_ptr(), X.data_ptr(), X.data_ptr(), 100,
                                 100, 16, 16, 16, C.byref(o), None)
    assert st == 1


def _push(lib, X, send_rows, seg_begin, seg_rows, dsts, dst_row, ld_dst, F, elem, mover, ctas=0, warps=0):
    W = len(seg_rows)
    o = _lib.HaloOpts()
    o.struct_size = C.sizeof(_lib.HaloOpts)
    o.mover, o.ctas, o.warps_per_cta, o.first_peer = mover, ctas, warps, 1 % W
    return lib.gnn_halo_push(X.data_ptr(), X.stride(0), F, elem, None if send_rows is None else send_rows.data_ptr(),
                             (C.c_int64 * W)(*seg_begin), (C.c_int64 * W)(*seg_rows),
                             (C.c_void_p * W)(*[d.data_ptr() for d in dsts]), (C.c_int64 * W)(*dst_row), ld_dst, W,
                             C.byref(o), torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("mover,ctas,warps", [(1, 0, 0), (2, 0, 0), (2, 7, 4), (2, 148, 1), (1, 16, 32)])
@pytest.mark.parametrize("dtype,F", [(torch.float32, 128), (torch.bfloat16, 128), (torch.float32, 64), (torch.float32, 604)])
def test_halo_push_movers_bit_exact(lib, mover, ctas, warps, dtype, F):
    """Rows land where the plan says, for ragged segments (empty, 1 row, not a multiple of the stage)."""
    n, W = 50_000, 4
    X = torch.randn(n, F, device=DEV).to(dtype)
    g = torch.Generator().manual_seed(5)
    seg_rows = [0, 1, 4099, 33_333]
    send = torch.randint(0, n, (sum(seg_rows) + 10,), generator=g, dtype=torch.int32).to(DEV)
    seg_begin = [3, 3, 4, 4103]
    dsts = [torch.full((40_000, F), -7.0, device=DEV).to(dtype) for _ in range(W)]
    dst_row = [0, 11, 5, 100]
    elem = X.element_size()
    _lib.check(_push(lib, X, send, seg_begin, seg_rows, dsts, dst_row, F, F, elem, mover, ctas, warps), "gnn_halo_push")
    torch.cuda.synchronize()
    for q in range(W):
        want = torch.full_like(dsts[q], -7.0)
        want[dst_row[q]:dst_row[q] + seg_rows[q]] = X[send[seg_begin[q]:seg_begin[q] + seg_rows[q]].long()]
        assert torch.equal(dsts[q], want), q
    # identity segments (send_rows == NULL): the contiguous reverse exchange of the backward
    dsts = [torch.zeros((40_000, F), device=DEV).to(dtype) for _ in range(W)]
    _lib.check(_push(lib, X, None, [0, 10, 20, 5000], seg_rows, dsts, dst_row, F, F, elem, mover, ctas, warps), "push")
    torch.cuda.synchronize()
    for q, b in enumerate([0, 10, 20, 5000]):
        assert torch.equal(dsts[q][dst_row[q]:dst_row[q] + seg_rows[q]], X[b:b + seg_rows[q]])


def test_halo_push_tma_refuses_what_it_cannot_move(lib):
    X = torch.randn(100, 130, device=DEV)  # 520-byte rows: not a 16-byte multiple
    d = torch.zeros(100, 130, device=DEV)
    send = torch.arange(100, dtype=torch.int32, device=DEV)
    assert _push(lib, X, send, [0], [100], [d], [0], 130, 130, 4, mover=2) == 3  # GNN_ERR_UNSUPPORTED
    _lib.check(_push(lib, X, send, [0], [100], [d], [0], 130, 130, 4, mover=0), "auto falls back to vector stores")
    torch.cuda.synchronize()
    assert torch.equal(d, X)


def test_peer_flags_signal_then_wait_and_timeout(lib):
    """Signal and wait on ONE stream (the value is already there when the wait starts); a wait for a value
    nobody signals must give up after its timeout and report the late slot instead of hanging."""
    W = 4
    flags = [torch.zeros(32, dtype=torch.int32, device=DEV) for _ in range(W)]
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    for me in range(W):  # every "rank" signals every other rank's array in its own slot
        ptrs = (C.c_void_p * W)(*[f.data_ptr() for f in flags])
        _lib.check(lib.gnn_peer_signal(ptrs, W, me, me, 5, st), "gnn_peer_signal")
    for me in range(W):
        _lib.check(lib.gnn_peer_wait(flags[me].data_ptr(), W, me, 5, status.data_ptr(), 2000, st), "gnn_peer_wait")
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    for me in range(W):
        want = [5] * W
        want[me] = 0
        assert flags[me][:W].tolist() == want
    _lib.check(lib.gnn_peer_wait(flags[0].data_ptr(), W, 0, 6, status.data_ptr(), 50, st), "gnn_peer_wait")
    torch.cuda.synchronize()
    assert int(status.item()) in (2, 3, 4)  # 1 + a late slot


@pytest.mark.parametrize("waves,c0", [(1, 1), (1, 0), (4, None), (4, 2)])
def test_single_process_emulation_of_a_two_rank_step(lib, waves, c0):
    """Both ranks' plans built in one process (the plan's collectives replaced by hand), the exchange done
    with the real mover into the other rank's halo buffer on the same GPU, the consumers run through
    spmm_ex in the scheduled order: result == the 1-GPU SpMM of the whole graph."""
    n, F, world = 6000, 128, 2
    rowptr, col, val = _graph(n, n, 12, seed=11, long_row=(10, 4000))
    rng = np.random.default_rng(2)
    X = rng.standard_normal((n, F)).astype(np.float32)
    bounds = [0, 2900, n]
    blocks = []
    for r in range(world):
        lo, hi = bounds[r], bounds[r + 1]
        blocks.append((torch.from_numpy(rowptr[lo:hi + 1] - rowptr[lo]), torch.from_numpy(col[rowptr[lo]:rowptr[hi]]),
                       torch.from_numpy(val[rowptr[lo]:rowptr[hi]])))
    plans = _two_rank_plans(blocks, bounds, waves, c0, F)
    g_full = CSRGraph(cuda(rowptr), cuda(col), cuda(val), n, n)
    ref = Fn.spmm_raw(g_full, cuda(X))
    Xs = [cuda(X[bounds[r]:bounds[r + 1]]) for r in range(world)]
    halos = [torch.full((max(p.n_halo, 1), F), float("nan"), device=DEV) for p in plans]
    outs = []
    for r, plan in enumerate(plans):  # exchange: rank r pushes, wave by wave, into rank 1-r's halo
        q = 1 - r
        send = plan.send_rows.to(DEV)
        for w in range(plan.waves):
            b = sum(plan.send_wave_counts[q][:w])
            seg_begin, seg_rows, dst_row = [0, 0], [0, 0], [0, 0]
            seg_begin[q] = plan.send_off()[q] + b
            seg_rows[q] = plan.send_wave_counts[q][w]
            dst_row[q] = plan.dst_off[q] + b
            _lib.check(_push(lib, Xs[r], send, seg_begin, seg_rows, [halos[0], halos[1]], dst_row, F, F, 4, 2), "push")
    torch.cuda.synchronize()
    for r, plan in enumerate(plans):
        assert torch.equal(halos[r][:plan.n_halo], cuda(X)[plan.halo_ids.to(DEV)])
        op = PartitionedSpmm.__new__(PartitionedSpmm)  # consumers only: no process group in this test
        op.plan, op.dev, op._cons = plan, torch.device(DEV), {}
        Y = torch.full((plan.n_local, F), float("nan"), device=DEV)
        if plan.p1 is not None:
            op._run(plan.p1, Xs[r], halos[r], Y)
        for c in plan.p2:
            if c is not None:
                op._run(c, Xs[r], halos[r], Y)
        outs.append(Y)
    got = torch.cat(outs, 0)
    assert not torch.isnan(got).any()
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-5


def _two_rank_plans(blocks, bounds, waves, c0, F):
    """build_halo_plan for 2 ranks without a process group: the three all_to_all_single calls of the plan
    are replayed by running both ranks in lock-step threads over an in-process exchange."""
    import threading
    import torch.distributed as dist
    world = 2
    box, bar = {}, threading.Barrier(world)
    tls = threading.local()

    def fake_all_to_all(out, inp, output_split_sizes=None, input_split_sizes=None, group=None):
        r = tls.rank
        n_in = inp.shape[0]
        ins = input_split_sizes or [n_in // world] * world
        outs_ = output_split_sizes or [out.shape[0] // world] * world
        off = np.concatenate([[0], np.cumsum(ins)])
        box[r] = [inp[off[q]:off[q + 1]].clone() for q in range(world)]
        bar.wait()
        pos = 0
        for q in range(world):
            out[pos:pos + outs_[q]] = box[q][r]
            pos += outs_[q]
        bar.wait()

    plans, errs = [None] * world, []
    orig = dist.all_to_all_single
    dist.all_to_all_single = fake_all_to_all

    def run(r):
        tls.rank = r
        try:
            plans[r] = build_halo_plan(*blocks[r], bounds, r, world, waves=waves, two_pass_chunks=c0, F=F)
        except Exception as e:  # pragma: no cover
            errs.append(e)
            bar.abort()

    try:
        ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    finally:
        dist.all_to_all_single = orig
    assert not errs, errs
    return plans


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_two_rank_partitioned_spmm_all_transports_full_check():
    """Two real ranks over NVLink (a 2-rank cross-check of every transport): tools/spmm_dist.py on an
    8 M-row graph compares EVERY row of the partitioned forward with a 1-GPU SpMM and the partitioned backward with
    a 1-GPU transposed SpMM, for the single-launch TMA mover (p2p), the copy-engine transport (ce) and the NCCL
    baseline, fixed and automatic schedules.  The command is the one the round-2 hardware runs used
    (tools/run_r2_dist2b.sh / run_r2_dist2c.sh)."""
    import json
    import os
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tools", "spmm_dist.py"), "--nodes", "8000000", "--p-local", "0.8",
           "--window", "200000", "--scatter", "--full-check", "--backward", "--steps", "3", "--warmup", "1",
           "--transports", "p2p", "ce", "nccl", "--configs", "4:auto:tma:0:1", "auto:auto:tma:-1:-1"]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 6, out.stdout[-2000:]
    for d in lines:
        tag = (d["transport"], d["waves"], d["two_pass_chunks"])
        assert d["ok"] and d["backward_ok"], tag
        assert d["full_check_max_rel_err"] <= 1e-5 and d["full_check_backward_max_rel_err"] <= 1e-5, tag
        assert d["adjoint_rel_err"] <= 1e-6, tag
