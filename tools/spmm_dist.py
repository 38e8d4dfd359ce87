#!/usr/bin/env python
"""Partitioned SpMM driver (torchrun, one rank per GPU): papers100M-shaped power-law graph,
1-D row partition balanced by nnz, wave-pipelined halo exchange over NVLink.  JSON lines on rank 0.
    torchrun --nproc-per-node N tools/spmm_dist.py [--nodes ...] [--F 128] --configs "K:c0:mover:ctas:warps[:dedicated]" ...
Every configuration is checked against a float64 recomputation of sampled rows (X is a hash of the global
row id, so any rank can recompute any row) and, with --backward, through the adjoint identity
<dY, A·X> = <Aᵀ·dY, X> reduced over the ranks."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from graphneuralnetwork_b200 import _lib, functional as Fn, synthetic as S  # noqa: E402
from graphneuralnetwork_b200.graph import _p, _stream_ptr  # noqa: E402
from graphneuralnetwork_b200.partition import PartitionedSpmm, balanced_bounds, build_halo_plan, select_rows  # noqa: E402


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
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--p-local", type=float, default=0.0)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--skew", type=float, default=3.0)
    ap.add_argument("--scatter", action="store_true")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"])
    ap.add_argument("--transports", nargs="+", default=["p2p"])
    ap.add_argument("--configs", nargs="+", default=["4:auto:tma:0:0"],
                    help="waves:two_pass_chunks(auto|int):mover(tma|vector):ctas:warps[:dedicated_sms[:sep]] (sep = one mover "
                         "launch per wave instead of the single-launch wave mover)")
    ap.add_argument("--backward", action="store_true")
    ap.add_argument("--phases", action="store_true")
    ap.add_argument("--consumers", action="store_true", help="time every consumer launch of the step alone")
    ap.add_argument("--full-check", action="store_true", help="small graphs: compare every row with a 1-GPU SpMM")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if a.dtype == "f32" else torch.bfloat16
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    csr, bounds, nnz_total = build_rank_block(a.n, a.deg, rank, world, dev, skew=a.skew, p_local=a.p_local,
                                              window=a.window, scatter=a.scatter)
    lo, hi = bounds[rank], bounds[rank + 1]
    n_loc = hi - lo
    X = S.hashed_feature_block(lo, hi, a.F, dev, dtype)
    rows = S.spmm_check_rows(csr, 4096, 17 + rank)
    ref_rows = S.spmm_sampled_reference(csr, rows, a.F, dtype)  # the checker sees the same (rounded) inputs
    ref_scale = float(ref_rows.abs().max().item())

    def reduce_max(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def reduce_sum(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t)
        return t.item()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return reduce_max(t0.elapsed_time(t1) / steps)

    Y = torch.empty(n_loc, a.F, device=dev, dtype=dtype)
    configs = a.configs if world > 1 else ["1:auto:tma:0:0"]
    plans, plan = {}, None
    for transport in (a.transports if world > 1 else ["none"]):
        for cfg in configs:
            parts = cfg.split(":")
            K = None if parts[0] == "auto" else int(parts[0])  # auto: (K, c0) by the cross-rank byte model
            c0 = None if parts[1] == "auto" else int(parts[1])
            mover, ctas, warps = parts[2], int(parts[3]), int(parts[4])
            dedicated = int(parts[5]) if len(parts) > 5 else 0
            fused = (parts[6] != "sep") if len(parts) > 6 else True
            inter = int(parts[7]) if len(parts) > 7 else 0
            if transport == "nccl":
                K, c0 = 1, (1 if c0 is None else min(c0, 1))
            want_K = K
            key = (K, c0)
            if key not in plans:
                plans.clear()
                plan = None
                torch.cuda.empty_cache()
                plans[key] = build_halo_plan(csr.rowptr, csr.col, csr.val, bounds, rank, world, waves=K,
                                             two_pass_chunks=c0, F=a.F, elem_size=X.element_size())
            plan = plans[key]
            op = PartitionedSpmm(plan, a.F, dev, transport=transport, dtype=dtype, mover=mover, mover_ctas=ctas,
                                 mover_warps=warps, dedicated_sms=dedicated, fused_signal=fused, interleave=inter)
            ms = timed(lambda: op.forward(X, out=Y), a.steps, a.warmup)
            op.check_status()
            err = reduce_max(float((Y[rows].double() - ref_rows).abs().max().item()) / max(ref_scale, 1e-30))
            stats = torch.tensor([plan.n_halo, int(plan.send_rows.numel()), plan.model.get("rows_interior", 0),
                                  plan.two_pass_chunks], dtype=torch.int64, device=dev)
            allstats = [torch.empty_like(stats) for _ in range(world)]
            if world > 1:
                dist.all_gather(allstats, stats)
            else:
                allstats = [stats]
            line = {"bench": "spmm_partitioned", "world": world, "n": a.n, "nnz": nnz_total, "F": a.F, "dtype": a.dtype,
                    "transport": transport, "waves": plan.waves, "waves_requested": "auto" if want_K is None else want_K, "two_pass_chunks": [int(s[3]) for s in allstats],
                    "mover": mover, "fused_signal": bool(op._wave_table is not None) if world > 1 else None, "interleave": inter, "mover_ctas": ctas, "mover_warps": warps, "dedicated_sms": dedicated,
                    "ms": ms, "edges_per_s": nnz_total / ms * 1e3, "max_rel_err_sampled_rows": err, "ok": err < tol,
                    "p_local": a.p_local, "window": a.window,
                    "halo_gb_recv_max": max(int(s[0]) for s in allstats) * a.F * X.element_size() / 1e9,
                    "send_gb_max": max(int(s[1]) for s in allstats) * a.F * X.element_size() / 1e9,
                    "interior_row_frac": sum(int(s[2]) for s in allstats) / a.n,
                    "mover_geometry": [op.mover_ctas, op.mover_warps],
                    "model_ms": {k: (round(v, 2) if isinstance(v, float) else v) for k, v in plan.model.items()
                                 if k.startswith("c0=") or k in ("exchange_ms", "auto_schedule_ms_sum_over_ranks")}}
            if a.backward:
                dY = S.hashed_feature_block(lo, hi, a.F, dev, dtype, salt=977)
                dX = torch.empty_like(Y)
                line["backward_ms"] = timed(lambda: op.backward(dY, out=dX), max(a.steps // 2, 2), 1)
                op.check_status()
                lhs = reduce_sum(float((dY.double() * Y.double()).sum().item()))
                rhs = reduce_sum(float((dX.double() * X.double()).sum().item()))
                norm = reduce_sum(float((dY.double().abs() * Y.double().abs()).sum().item()))
                line["adjoint_rel_err"] = abs(lhs - rhs) / max(norm, 1e-30)
                line["backward_ok"] = line["adjoint_rel_err"] < (1e-6 if dtype == torch.float32 else 1e-2)
                del dY, dX
            if a.consumers and world > 1:
                # each consumer launch alone (halo already in place): which pass runs below the 1-GPU SpMM rate?
                rows_c = []
                for name, c in [("P1", plan.p1)] + [(f"P2[{w}]", c) for w, c in enumerate(plan.p2)]:
                    if c is None:
                        continue
                    ms_c = timed(lambda: op._run(c, X, op.halo, Y), 3, 1)
                    acc = 2 if c.mode == "remote" else 1
                    gb = (c.nnz * (8 + a.F * X.element_size()) + c.n_rows * a.F * X.element_size() * acc) / 1e9
                    rows_c.append({"pass": name, "mode": c.mode, "rows": c.n_rows, "nnz": c.nnz, "ms": round(ms_c, 3),
                                   "gather_model_gbs": round(gb / ms_c * 1e3, 1)})
                line["consumers_alone"] = rows_c
                op.forward(X, out=Y)  # leave Y complete again
            if a.phases and world > 1:
                t_x = timed(lambda: op.exchange_only(X), 4, 1)
                line["exchange_alone_ms"] = t_x
                line["exchange_alone_gbs"] = line["halo_gb_recv_max"] / t_x * 1e3
            if a.full_check:
                full = S.powerlaw_csr(a.n, a.deg, seed=0, device=dev, skew=a.skew, p_local=a.p_local, window=a.window,
                                      scatter_hubs=a.scatter)
                Xf = S.hashed_feature_block(0, a.n, a.F, dev, dtype)
                ref = Fn.spmm_raw(full, Xf)[lo:hi]
                line["full_check_max_rel_err"] = reduce_max(float((Y.float() - ref.float()).abs().max().item()) /
                                                            max(float(ref.float().abs().max().item()), 1e-30))
                if a.backward:
                    dYf = S.hashed_feature_block(0, a.n, a.F, dev, dtype, salt=977)
                    refb = Fn.spmm_raw(full.transpose(), dYf)[lo:hi]
                    dX = op.backward(dYf[lo:hi].contiguous())
                    line["full_check_backward_max_rel_err"] = reduce_max(
                        float((dX.float() - refb.float()).abs().max().item()) / max(float(refb.float().abs().max().item()), 1e-30))
                del full, Xf
            if rank == 0:
                print(json.dumps(line), flush=True)
            op.close()
            del op
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
