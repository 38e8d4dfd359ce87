"""Host logic of the 1-D row partition + halo plan, world_size 2 and 3 over gloo on CPU.
The exchange is replayed with torch/gloo collectives and the aggregation with a float64
scipy product (test-side stand-ins; the product path uses the CUDA kernels)."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphneuralnetwork_b200.partition import balanced_bounds, build_halo_plan, split_columns


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_graph(n, seed):
    rng = np.random.default_rng(seed)
    deg = rng.poisson(6, size=n)
    deg[::11] = 0
    deg[5] = 300  # a hub row
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    # skewed targets: low ids are hubs, so many ranks request the same rows (exercises dedup)
    col = np.minimum((n * rng.random(rowptr[-1]) ** 3).astype(np.int64), n - 1)
    val = rng.standard_normal(rowptr[-1]).astype(np.float32)
    X = rng.standard_normal((n, 5)).astype(np.float32)
    return rowptr, col, val, X


def _worker(rank, world, port, n, seed, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rowptr, col, val, X = _global_graph(n, seed)
        bounds = balanced_bounds(torch.from_numpy(rowptr), world)
        lo, hi = bounds[rank], bounds[rank + 1]
        rp = torch.from_numpy(rowptr[lo:hi + 1] - rowptr[lo])
        c = torch.from_numpy(col[rowptr[lo]:rowptr[hi]])
        v = torch.from_numpy(val[rowptr[lo]:rowptr[hi]])
        plan = build_halo_plan(rp, c, v, bounds, rank, world)
        # invariants of the plan
        assert plan.n_local == hi - lo
        h = plan.halo_ids.numpy()
        assert np.all(np.diff(h) > 0)                       # sorted, de-duplicated
        assert not np.any((h >= lo) & (h < hi))             # never asks for its own rows
        assert sum(plan.recv_counts) == len(h) and plan.recv_counts[rank] == 0
        assert plan.send_counts[rank] == 0 and int(plan.send_rows.numel()) == sum(plan.send_counts)
        assert plan.rowptr_loc[-1] + plan.rowptr_rem[-1] == rp[-1]
        # replay the exchange
        Xl = torch.from_numpy(X[lo:hi])
        send = Xl[plan.send_rows.long()]
        halo = torch.empty((len(h), X.shape[1]))
        dist.all_to_all_single(halo, send, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts)
        assert np.array_equal(halo.numpy(), X[h])           # every halo slot holds the right global row
        # dst_off: owner q's segment for requester p starts where p's halo lists q's rows
        offs = torch.tensor(plan.dst_off)
        gathered = [torch.empty_like(offs) for _ in range(world)]
        dist.all_gather(gathered, offs)
        my_off = np.concatenate([[0], np.cumsum(plan.recv_counts)])
        for q in range(world):
            assert int(gathered[q][rank]) == int(my_off[q])
        # aggregate: local columns + remote columns == the global product's rows
        A_loc = sp.csr_matrix((plan.val_loc.numpy().astype(np.float64), plan.col_loc.numpy(), plan.rowptr_loc.numpy()),
                              shape=(hi - lo, hi - lo))
        A_rem = sp.csr_matrix((plan.val_rem.numpy().astype(np.float64), plan.col_rem.numpy(), plan.rowptr_rem.numpy()),
                              shape=(hi - lo, max(len(h), 1)))
        Y = A_loc @ X[lo:hi].astype(np.float64) + (A_rem @ halo.numpy().astype(np.float64) if len(h) else 0)
        A = sp.csr_matrix((val.astype(np.float64), col, rowptr), shape=(n, n))
        ref = (A @ X.astype(np.float64))[lo:hi]
        err = float(np.abs(Y - ref).max())
        out_q.put((rank, err, len(h), bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_plan_reproduces_global_spmm(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 400, 7, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, n_halo, bounds in results:
        assert err < 1e-9, (rank, err)
        assert n_halo > 0 and bounds[0] == 0 and bounds[-1] == 400


def test_balanced_bounds_by_nnz():
    rowptr, _, _, _ = _global_graph(1000, 3)
    for world in (1, 2, 4, 8):
        b = balanced_bounds(torch.from_numpy(rowptr), world)
        assert b[0] == 0 and b[-1] == 1000 and all(x <= y for x, y in zip(b, b[1:]))
        nnz = [rowptr[b[i + 1]] - rowptr[b[i]] for i in range(world)]
        assert max(nnz) - min(nnz) <= 2 * 300 + 20  # within one hub row of perfect balance


def test_split_columns_preserves_order():
    rowptr = torch.tensor([0, 3, 3, 6])
    col = torch.tensor([0, 5, 1, 4, 2, 5])
    val = torch.arange(6, dtype=torch.float32)
    rl, cl, vl, rr, cr, vr = split_columns(rowptr, col, val, 0, 3)
    assert rl.tolist() == [0, 2, 2, 3] and cl.tolist() == [0, 1, 2] and vl.tolist() == [0., 2., 4.]
    assert rr.tolist() == [0, 1, 1, 3] and cr.tolist() == [5, 4, 5] and vr.tolist() == [1., 3., 5.]
