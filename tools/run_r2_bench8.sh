#!/bin/bash
# round 2: the driver's N=8 launch of bench.py after the exchange-bound rule (single-pass chunks + 32 mover CTAs)
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 \
  bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2_bench_8gpu_v2.json 2> $O/r2_bench_8gpu_v2.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_8gpu_v2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], json.dumps(d.get("strong_scaling"))[:1200])
for g in ("random", "locality"):
    p = d["partitioned_spmm"][g]; print(g, p["ms"], p["schedule"], p["transport"][:90])
PY
tail -3 $O/r2_bench_8gpu_v2.err
