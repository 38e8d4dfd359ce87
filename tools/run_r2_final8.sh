#!/bin/bash
# Round 2, final 8-GPU validation: the driver's bench launch at N=8 and N=4, then the partitioned SpMM with the
# automatic schedule, interleaved segment dealing, and the bf16 variant (bf16 rows on the wire)
O=gpurun_out
mkdir -p $O
tr() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 "${@:4}"; }
tr 600 8 29561 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2_bench_8gpu.json 2> $O/r2_bench_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_8gpu.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d.get("e2e", {}).get("value"), json.dumps(d.get("strong_scaling"))[:1500])
except Exception as e:
    print("bench8 parse failed", e)
PY
tail -3 $O/r2_bench_8gpu.err
tr 600 4 29562 bench.py --gpus 4 --steps 20 --warmup 3 > $O/r2_bench_4gpu.json 2> $O/r2_bench_4gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_4gpu.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, json.dumps(d.get("strong_scaling"))[:1500])
except Exception as e:
    print("bench4 parse failed", e)
PY
run() { tr $1 8 $2 tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-330; }
run 300 29563 r2h_loc --p-local 0.8 --window 2000000 --scatter --steps 8 --warmup 2 --transports p2p \
  --configs auto:auto:tma:-1:-1 4:4:tma:48:4:0:fused:16 4:4:tma:48:4:0:fused:64 4:4:tma:48:4:0:fused:256
run 300 29564 r2h_loc_bf16 --p-local 0.8 --window 2000000 --scatter --steps 8 --warmup 2 --transports p2p --dtype bf16 \
  --configs auto:auto:tma:-1:-1 4:4:tma:48:4:0:fused:0 2:2:tma:48:4:0:fused:0 4:0:tma:32:4:0:fused:0
run 300 29565 r2h_rand_bf16 --scatter --steps 6 --warmup 2 --transports p2p --dtype bf16 \
  --configs auto:auto:tma:-1:-1 8:0:tma:32:4:0:fused:0
