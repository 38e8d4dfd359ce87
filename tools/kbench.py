#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events, L2 flushed between reps) -> JSON lines.
    python tools/kbench.py sage|spmm_reddit|spmm_papers|gat [options]
Used for tuning sweeps on the GPU box; bench.py is the contract benchmark."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from graphneuralnetwork_b200 import _lib, functional as Fn, synthetic as S  # noqa: E402
from graphneuralnetwork_b200.graph import CSRGraph  # noqa: E402

DEV = torch.device("cuda")
PEAK = 6451.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    _flush.zero_()


def timeit(fn, reps=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(reps):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def bench_sage(args):
    n, F = 232_965, args.F
    dt = torch.bfloat16 if args.bf16 else torch.float32
    table = Fn.pad_table(torch.randn(n, F, device=DEV).to(dt))
    es = 2 if args.bf16 else 4
    for n_src, fanout in ((25600, 10), (1024, 25)):
        idx = torch.randint(0, n, (n_src * fanout,), device=DEV, dtype=torch.int32)
        out = Fn._padded_empty(n_src, F, dt, DEV)
        B = n_src * fanout * (4 + F * es) + n_src * F * es
        for knobs in args.knobs:
            for k, v in knobs.items():
                _lib.set_tuning(k, v)
            med, best = timeit(lambda: Fn.gather_reduce_raw(table, idx, n_src, fanout, "mean", out=out), reps=args.reps)
            emit(bench="sage", F=F, bf16=args.bf16, n_src=n_src, fanout=fanout, knobs=knobs, ms=med, ms_best=best,
                 gbs=B / med / 1e6, frac=B / med / 1e6 / PEAK, edges_per_s=n_src * fanout / med * 1e3)


def bench_spmm(args, n, mean_deg, tag):
    csr = S.powerlaw_csr(n, mean_deg, seed=0, device=DEV, exponent=args.exponent, skew=args.skew,
                         max_degree=args.max_degree)
    deg = csr.rowptr[1:] - csr.rowptr[:-1]
    emit(bench=tag + "_graph", n=n, nnz=csr.nnz, max_deg=int(deg.max()), long_rows=int(csr.long_rows().numel()))
    for F in args.Fs:
        for dt in ([torch.bfloat16] if args.only_bf16 else [torch.float32, torch.bfloat16] if args.bf16 else [torch.float32]):
            # 16-byte rows, as bench.py and the drop-in layers allocate them (602 -> ld 604 / 608)
            X = Fn.pad_table(torch.randn(n, F, device=DEV).to(dt))
            Y = Fn._padded_empty(n, F, dt, DEV)
            es = X.element_size()
            B = csr.nnz * 8 + csr.nnz * F * es + n * F * es + (n + 1) * 8
            comp = csr.nnz * 8 + 2 * n * F * es + (n + 1) * 8
            for planned in ([True, False] if args.both else [True]):
                for knobs in args.knobs:
                    for k, v in knobs.items():
                        _lib.set_tuning(k, v)
                    med, best = timeit(lambda: Fn.spmm_raw(csr, X, out=Y, planned=planned), reps=args.reps)
                    emit(bench=tag, F=F, dtype=str(dt), planned=planned, knobs=knobs, rows_per_team=csr.rows_per_team(),
                         ms=med, ms_best=best, gather_gbs=B / med / 1e6, gather_frac=B / med / 1e6 / PEAK,
                         compulsory_gbs=comp / med / 1e6, edges_per_s=csr.nnz / med * 1e3, X_mb=n * F * es / 1e6)
            del X, Y


def bench_gat(args):
    for n, mean_deg, tag in ((2708, 4.9, "cora"), (3025, 730, "acm_dense"), (232_965, 100, "reddit_d100"),
                             (232_965, 492, "reddit_full")):
        if (args.graph and tag != args.graph) or (tag == "reddit_full" and args.graph != tag):
            continue
        # skew=1: uniform targets, so in-degrees stay moderate like in the symmetric adjacencies
        # GAT/HAN are given (the transposed graph the backward walks has no 100k-edge rows)
        csr = S.powerlaw_csr(n, mean_deg, seed=0, device=DEV, with_values=False, max_degree=min(n - 1, 20000),
                             skew=1.0)
        H, Fp = 8, 8
        fdt = torch.bfloat16 if args.bf16 else torch.float32
        Wh = torch.randn(n, H * Fp, device=DEV).to(fdt).requires_grad_(True)
        s = torch.randn(n, H, device=DEV, requires_grad=True)
        t = torch.randn(n, H, device=DEV, requires_grad=True)
        for knobs in args.knobs:
            for k, v in knobs.items():
                _lib.set_tuning(k, v)
            bench_gat_one(args, csr, Wh, s, t, H, Fp, n, tag, fdt, knobs)


def bench_gat_one(args, csr, Wh, s, t, H, Fp, n, tag, fdt, knobs):
    if True:
        med, best = timeit(lambda: Fn.gat_fwd_raw(csr, Wh.detach(), s.detach(), t.detach(), H, Fp, 0.2, elu=1), reps=args.reps)
        es = Wh.element_size()
        B = csr.nnz * (4 + H * Fp * es + H * 4) + n * (H * 4 + H * Fp * es) + (n + 1) * 8
        emit(bench="gat_fwd", graph=tag, dtype=str(fdt), knobs=knobs, n=n, nnz=csr.nnz, ms=med, ms_best=best, gather_gbs=B / med / 1e6,
             edges_per_s=csr.nnz / med * 1e3)
        csr.transpose()
        out = Fn.gat_aggregate(csr, Wh, s, t, H, Fp, 0.2)
        g = torch.randn_like(out)
        med, best = timeit(lambda: torch.autograd.grad(out, (Wh, s, t), g, retain_graph=True), reps=args.reps)
        emit(bench="gat_bwd", graph=tag, dtype=str(fdt), knobs=knobs, n=n, nnz=csr.nnz, ms=med, ms_best=best, edges_per_s=csr.nnz / med * 1e3)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--F", type=int, default=602)
    ap.add_argument("--Fs", type=int, nargs="+", default=[16, 64, 128, 602])
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--only-bf16", action="store_true")
    ap.add_argument("--both", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--deg", type=float, default=0)
    ap.add_argument("--exponent", type=float, default=2.5)
    ap.add_argument("--skew", type=float, default=3.0)
    ap.add_argument("--max-degree", type=int, default=1 << 20)
    ap.add_argument("--knobs", type=json.loads, default=[{}])
    ap.add_argument("--graph", default="")
    a = ap.parse_args()
    if a.what == "sage":
        bench_sage(a)
    elif a.what == "spmm_reddit":
        bench_spmm(a, a.n or 232_965, a.deg or 492.0, "spmm_reddit")
    elif a.what == "spmm_papers":
        bench_spmm(a, a.n or 111_059_956, a.deg or 13.55, "spmm_papers")
    elif a.what == "gat":
        bench_gat(a)
