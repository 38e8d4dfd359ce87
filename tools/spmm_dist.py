#!/usr/bin/env python
"""Partitioned SpMM driver (torchrun, one rank per GPU): papers100M-shaped power-law graph,
1-D row partition balanced by nnz, halo exchange over NVLink.  JSON lines on rank 0.
    torchrun --nproc-per-node N tools/spmm_dist.py [--n ...] [--deg ...] [--F 128] [--check]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from graphneuralnetwork_b200 import _lib, functional as Fn, synthetic as S  # noqa: E402
from graphneuralnetwork_b200.graph import _p, _stream_ptr  # noqa: E402
from graphneuralnetwork_b200.partition import PartitionedSpmm, balanced_bounds, build_halo_plan  # noqa: E402


def build_rank_block(n, deg, rank, world, dev, seed=0, exponent=2.5, skew=3.0, p_local=0.0, window=0,
                     max_degree=1 << 20, scatter=False):
    lib = _lib.load()
    deg_all = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.check(lib.gnn_synth_powerlaw_degrees(n, 0, float(deg), float(exponent), int(max_degree), seed, _p(deg_all),
                                              _stream_ptr()), "degrees")
    rowptr_g = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg_all, 0, out=rowptr_g[1:])
    bounds = balanced_bounds(rowptr_g, world)
    nnz_total = int(rowptr_g[-1].item())
    del rowptr_g
    lo, hi = bounds[rank], bounds[rank + 1]
    csr = S.powerlaw_csr(hi - lo, deg, n_cols=n, row_offset=lo, exponent=exponent, skew=skew, max_degree=max_degree,
                         seed=seed, device=dev, deg_all=deg_all, p_local=p_local, window=window, scatter_hubs=scatter)
    del deg_all
    return csr, bounds, nnz_total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", dest="n", type=int, default=S.PAPERS100M["n"])
    ap.add_argument("--deg", type=float, default=13.55)
    ap.add_argument("--F", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--p-local", type=float, default=0.0)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--skew", type=float, default=3.0)
    ap.add_argument("--transports", nargs="+", default=["p2p", "nccl"])
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--phases", action="store_true")
    ap.add_argument("--scatter", action="store_true")
    ap.add_argument("--halo-ctas", type=int, default=1)
    ap.add_argument("--halo-sched", type=int, default=0)
    ap.add_argument("--halo-unroll", type=int, nargs="+", default=[8])
    ap.add_argument("--overlap-only", action="store_true")
    ap.add_argument("--tma", action="store_true", help="experimental TMA mover for the p2p push (halo.tma)")
    ap.add_argument("--cross-check", action="store_true",
                    help="compare the last p2p result with the NCCL transport's (any graph size)")
    ap.add_argument("--dedicated", type=int, nargs="+", default=[0], help="PartitionedSpmm.dedicated values to sweep")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.set_tuning("halo.ctas_per_sm", a.halo_ctas)
    _lib.set_tuning("halo.schedule", a.halo_sched)
    _lib.set_tuning("halo.unroll", a.halo_unroll[0])
    csr, bounds, nnz_total = build_rank_block(a.n, a.deg, rank, world, dev, skew=a.skew, p_local=a.p_local, window=a.window,
                                                  scatter=a.scatter)
    plan = build_halo_plan(csr.rowptr, csr.col, csr.val, bounds, rank, world)
    n_loc = plan.n_local
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    X = torch.randn(n_loc, a.F, device=dev, generator=gen)
    stats = torch.tensor([plan.n_halo, csr.nnz, int(plan.rowptr_loc[-1].item()), int(plan.send_rows.numel())],
                         dtype=torch.int64, device=dev)
    allstats = [torch.empty_like(stats) for _ in range(world)]
    if world > 1:
        dist.all_gather(allstats, stats)
    else:
        allstats = [stats]
    del csr
    for transport in (a.transports if world > 1 else ["none"]):
        op = PartitionedSpmm(plan, a.F, dev, transport=transport)
        Y = torch.empty(n_loc, a.F, device=dev)
        # transports: p2p (fused NVLink push), nccl (library baseline), ce (experimental: local pack + copy engines)
        for overlap, dedicated, unroll in [(o, d, u) for u in (a.halo_unroll if transport == "p2p" else a.halo_unroll[:1])
                                           for d in (a.dedicated if transport == "p2p" else [0])
                                           for o in ([True] if (a.overlap_only or world == 1) else [True, False])]:
            if unroll != a.halo_unroll[0] and dedicated == 0:
                continue  # the shared-SM schedule is swept with the first unroll only
            op.dedicated = dedicated
            op.tma = bool(a.tma)
            _lib.set_tuning("halo.unroll", unroll)
            for _ in range(a.warmup):
                op.forward(X, out=Y, overlap=overlap)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(a.steps):
                op.forward(X, out=Y, overlap=overlap)
            t1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([t0.elapsed_time(t1) / a.steps], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                hal = [int(s[0]) for s in allstats]
                print(json.dumps({"bench": "spmm_partitioned", "world": world, "n": a.n, "nnz": nnz_total, "F": a.F,
                                  "transport": transport, "overlap": overlap, "dedicated_sms": dedicated, "tma": bool(a.tma), "halo_ctas": a.halo_ctas, "sched": a.halo_sched, "unroll": unroll, "ms": ms.item(),
                                  "edges_per_s": nnz_total / ms.item() * 1e3, "p_local": a.p_local, "window": a.window, "scatter": a.scatter, "skew": a.skew,
                                  "halo_rows_max": max(hal), "halo_rows_mean": sum(hal) / world,
                                  "halo_gb_recv_max": max(hal) * a.F * 4 / 1e9,
                                  "send_rows_max": max(int(s[3]) for s in allstats),
                                  "local_edge_frac": sum(int(s[2]) for s in allstats) / max(sum(int(s[1]) for s in allstats), 1),
                                  "bounds": bounds if world <= 8 else None}), flush=True)
        if world > 1 and a.phases:
            from graphneuralnetwork_b200.functional import spmm_raw

            def timed(fn, reps=5):
                fn()
                torch.cuda.synchronize()
                dist.barrier()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(reps):
                    fn()
                t1.record()
                torch.cuda.synchronize()
                m = torch.tensor([t0.elapsed_time(t1) / reps], device=dev, dtype=torch.float64)
                dist.all_reduce(m, op=dist.ReduceOp.MAX)
                return m.item()

            t_x = timed(lambda: op._exchange(X))
            t_l = timed(lambda: spmm_raw(op.A_loc, X, out=Y))
            halo = op.halo[0]
            t_r = timed(lambda: spmm_raw(op.A_rem, halo, out=Y, accumulate=True))
            if rank == 0:
                print(json.dumps({"phases": transport, "exchange_ms": t_x, "local_ms": t_l, "remote_ms": t_r,
                                  "exchange_gbs_recv": max(int(s[0]) for s in allstats) * a.F * 4 / t_x / 1e6}), flush=True)
        if a.cross_check and world > 1 and transport in ("p2p", "ce"):
            ref_op = PartitionedSpmm(plan, a.F, dev, transport="nccl")
            Yr = ref_op.forward(X, overlap=False)
            torch.cuda.synchronize()
            err = (Y - Yr).abs().max().item() / max(Yr.abs().max().item(), 1e-30)
            errs = torch.tensor([err], device=dev, dtype=torch.float64)
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(json.dumps({"cross_check": "p2p vs nccl transport", "max_rel_err": errs.item(),
                                  "ok": errs.item() < 1e-6}), flush=True)
            ref_op.close()
            del ref_op, Yr
        if a.check:
            # every rank rebuilds the FULL graph (small n only) and checks its own rows
            full = S.powerlaw_csr(a.n, a.deg, seed=0, device=dev, skew=a.skew, p_local=a.p_local, window=a.window,
                                  scatter_hubs=a.scatter)
            Xs = [torch.empty(bounds[q + 1] - bounds[q], a.F, device=dev) for q in range(world)]
            if world > 1:
                dist.all_gather(Xs, X)
            else:
                Xs = [X]
            ref = Fn.spmm_raw(full, torch.cat(Xs, 0))[bounds[rank]:bounds[rank + 1]]
            err = (Y - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
            errs = torch.tensor([err], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(errs, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(json.dumps({"check": transport, "max_rel_err": errs.item(), "ok": errs.item() < 1e-5}), flush=True)
            del full
        op.close()
        del op
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
