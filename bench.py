#!/usr/bin/env python
"""Benchmark of the message-passing hot path (BASELINE.json metric: aggregated edges/s +
HBM GB/s vs roofline) — one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Default workload `sage_reddit` = BASELINE.json configs[2]: GraphSAGE_Pytorch mean aggregator,
fanout (25,10), batch 1024, Reddit-shaped synthetic graph (232,965 nodes x 602 fp32 features,
resident in HBM), the config the metric is quoted on for one B200.  A step is one minibatch
through the aggregation hot path:
    hop-2 fused gather-mean  [25600 x 10 x 602]   (gnn_gather_reduce_f32, TMA ring)
    hop-1 fused gather-mean  [ 1024 x 25 x 602]
    layer-2 mean             [ 1024 x 25 x 128]   (identity block over the hidden tensor)
`value`   = sampled edges aggregated per second, kernels only, inputs resident in HBM.
`e2e`     = the same metric through the public drop-in API (GraphSage.forward_sampled): every
            step copies that step's sampled ids from pinned host memory, runs the whole model
            forward (aggregation + torch matmuls) and reads the logits back to the host.
`roofline`= the hop-2 kernel: algorithmic bytes n_src*fanout*(4+F*4)+n_src*F*4 per launch over
            its CUDA-event duration, against the measured HBM peak (MEASURED_PEAKS.json).
With N>1 every rank runs its own minibatches on its own replica of the table (SURVEY.md §8e:
"SAGE minibatch: replicas only") -> weak scaling, no data-path collective.
`--impl reference` times the reference's CPU arithmetic for the same step (the oracle port:
torch-CPU gather + GraphSage forward) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback, only if MEASURED_PEAKS.json is absent


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# workload: GraphSAGE_Pytorch mean aggregator on the Reddit-shaped graph
# ------------------------------------------------------------------------------------------
SAGE = dict(n=232_965, feats=602, batch=1024, fanout=(25, 10), hidden=(128, 41))
SAGE_EDGES = SAGE["batch"] * SAGE["fanout"][0] * (1 + SAGE["fanout"][1])  # 25,600 + 256,000 = 281,600


def sage_algorithmic_bytes(n_src, fanout, F, s=4):
    return n_src * fanout * (4 + F * s) + n_src * F * s


def sage_blocks_host(pool, seed):
    """`pool` minibatches of sampled ids (uniform: the throughput-run blocks of SURVEY.md §8d row 3),
    int32, in pinned host memory."""
    g = torch.Generator().manual_seed(seed)
    B, (f1, f2) = SAGE["batch"], SAGE["fanout"]
    out = []
    for _ in range(pool):
        blk = [torch.randint(0, SAGE["n"], (s,), generator=g, dtype=torch.int32) for s in (B, B * f1, B * f1 * f2)]
        out.append([b.pin_memory() if torch.cuda.is_available() else b for b in blk])
    return out


def run_sage_b200(args, rank, world, dev):
    from graphneuralnetwork_b200 import _lib, functional as Fn, layers
    lib = _lib.load()
    B, (f1, f2), F, H1 = SAGE["batch"], SAGE["fanout"], SAGE["feats"], SAGE["hidden"][0]
    gen = torch.Generator(device=dev).manual_seed(1234)
    table = Fn.pad_table(torch.randn(SAGE["n"], F, device=dev, generator=gen))  # 561 MB, >> 126 MB L2
    torch.manual_seed(0)
    model = layers.GraphSage(F, list(SAGE["hidden"]), list(SAGE["fanout"])).to(dev).eval()
    pool = 8  # distinct minibatches, cycled; each touches ~155k distinct rows (374 MB) of the table
    host_blocks = sage_blocks_host(pool, seed=100 + rank)
    dev_blocks = [[b.to(dev, non_blocking=True) for b in blk] for blk in host_blocks]
    hidden1 = torch.randn(B * f1, H1, device=dev, generator=gen)  # layer-1 output feeding the layer-2 mean
    out2 = Fn._padded_empty(B * f1, F, torch.float32, dev)
    out1 = Fn._padded_empty(B, F, torch.float32, dev)
    out0 = Fn._padded_empty(B, H1, torch.float32, dev)

    def hot_step(i, ev=None):
        blk = dev_blocks[i % pool]
        if ev is not None:
            ev[0].record()
        Fn.gather_reduce_raw(table, blk[2], B * f1, f2, "mean", out=out2)
        if ev is not None:
            ev[1].record()
        Fn.gather_reduce_raw(table, blk[1], B, f1, "mean", out=out1)
        Fn.gather_reduce_raw(hidden1, None, B, f1, "mean", out=out0)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- kernels-only leg -------------------------------------------------------------------
    for i in range(args.warmup):
        hot_step(i)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = lib.gnn_launch_count()
    with ClockSampler(dev.index or 0) as clocks:
        barrier()
        t0.record()
        for i in range(args.steps):
            hot_step(args.warmup + i, evs[i])
        t1.record()
        barrier()
    launches = lib.gnn_launch_count() - launches0
    ms_total = t0.elapsed_time(t1)
    k2_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))

    # ---- end-to-end leg through the public API ----------------------------------------------
    logits_host = torch.empty((B, SAGE["hidden"][1]), dtype=torch.float32).pin_memory()

    def e2e_step(i):
        hb = host_blocks[i % pool]
        ids = [b.to(dev, non_blocking=True) for b in hb]          # H2D: this step's sampled ids
        with torch.no_grad():
            logits = model.forward_sampled(table, ids)            # public drop-in API
        logits_host.copy_(logits, non_blocking=True)             # D2H: the step's result
        torch.cuda.current_stream().synchronize()
        return logits_host

    for i in range(args.warmup):
        e2e_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        e2e_step(args.warmup + i)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    if world > 1:
        t = torch.tensor([ms_total, e2e_ms, k2_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total, e2e_ms, k2_ms = t.tolist()

    peak, peak_src = hbm_peak()
    k2_bytes = sage_algorithmic_bytes(B * f1, f2, F)
    step_bytes = k2_bytes + sage_algorithmic_bytes(B, f1, F) + (B * f1 * H1 * 4 + B * H1 * 4)
    achieved = k2_bytes / (k2_ms * 1e-3) / 1e9
    h2d = sum(int(b.numel()) * b.element_size() for b in host_blocks[0])
    d2h = logits_host.numel() * 4
    res = {
        "metric": "aggregated_edges_per_sec",
        "value": world * args.steps * SAGE_EDGES / (ms_total * 1e-3),
        "unit": "edges/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "sage_reddit: GraphSAGE_Pytorch mean aggregator, fanout (25,10), batch 1024, "
                               "Reddit-shaped table 232,965 x 602 fp32 (BASELINE.json configs[2])",
                   "edges_per_step": SAGE_EDGES, "launches_per_step": 3, "minibatch_pool": pool,
                   "l2_policy": "inputs larger than L2: 561 MB table, ~374 MB distinct rows per step, pool of 8 "
                                "distinct minibatches cycled",
                   "parallelism": f"replicas x{world} (independent minibatches per rank, no collective)"},
        "hbm_gbs_step": step_bytes / (ms_total / args.steps * 1e-3) / 1e9,
        "roofline": {"bound": "hbm", "kernel": "sage_tma_kernel<float,1,SUM> hop-2 gather-mean [25600x10x602]",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": k2_bytes, "launch_ms": k2_ms,
                     "traffic": None},
        "e2e": {"value": world * args.steps * SAGE_EDGES / (e2e_ms * 1e-3), "unit": "edges/s",
                "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "graphneuralnetwork_b200.layers.GraphSage.forward_sampled(table, ids) "
                       "(aggregation kernels + torch matmuls, logits read back)"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    return res


def run_sage_cpu(args, steps, threads=None):
    """The reference's CPU arithmetic for one step (oracle port of GraphSAGE_Pytorch):
    feature gather (data_utils.py:64, vectorised) + GraphSage.forward (GraphSage.py:18-30)."""
    from oracle import sage as osage
    if threads:
        torch.set_num_threads(threads)
    F = SAGE["feats"]
    g = torch.Generator().manual_seed(1234)
    table = torch.randn(SAGE["n"], F, generator=g)
    torch.manual_seed(0)
    params = {}
    dims = [F] + list(SAGE["hidden"])
    for l in range(2):
        for name in ("weight", "aggregator.weight"):
            w = torch.empty(dims[l], dims[l + 1])
            torch.nn.init.xavier_uniform_(w)
            params[f"gcn.{l}.{name}"] = w
    blocks = sage_blocks_host(2, seed=100)

    def step(i):
        ids = [b.to(torch.int64) for b in blocks[i % 2]]
        feats = [osage.gather_features(table, b.numpy()) for b in ids]
        with torch.no_grad():
            return osage.graphsage_forward(feats, params, list(SAGE["fanout"]))

    step(0)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    dt = time.perf_counter() - t0
    return steps * SAGE_EDGES / dt, dt / steps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sage_reddit", choices=["sage_reddit"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 20))  # bounded sample: whole minibatches, ~0.2-0.5 s each on the host
        v, ms = run_sage_cpu(args, steps)
        line = {"impl": "reference", "metric": "aggregated_edges_per_sec", "value": v, "unit": "edges/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "sage_reddit: GraphSAGE_Pytorch mean aggregator, fanout (25,10), batch 1024, "
                                       "Reddit-shaped table 232,965 x 602 fp32 (BASELINE.json configs[2])",
                           "edges_per_step": SAGE_EDGES},
                "cpu_baseline": {"value": v, "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{steps} full minibatches: torch-CPU feature gather of the 3 id blocks + "
                                           "GraphSage forward (oracle/sage.py restating GraphSAGE_Pytorch); the Python "
                                           "reference itself cannot travel to the GPU box"},
                "e2e": {"value": v, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "host_cores": cores}
        print(json.dumps(line))
        return 0

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the hot path has no CPU fallback"}))
        return 1
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    res = run_sage_b200(args, rank, world, dev)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            v, ms = run_sage_cpu(args, 5)
            res["cpu_baseline"] = {"value": v, "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "ms_per_step": ms,
                                   "sample": "5 full minibatches of the same workload on the host: torch-CPU feature "
                                             "gather + GraphSage forward (oracle/sage.py)"}
        print(json.dumps(res))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
