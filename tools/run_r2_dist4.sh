#!/bin/bash
# 4 GPUs: the automatic schedule and the defaults at the one world size not measured yet (both graphs), plus two
# hand-picked schedules beside it.
set -x
O=gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $2 \
  tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-700; }
run 400 29561 r2m_loc4 --p-local 0.8 --window 2000000 --scatter --steps 6 --warmup 2 --transports p2p --configs auto:auto:tma:-1:-1 4:0:tma:32:4 4:4:tma:48:4 4:2:tma:48:4
run 300 29562 r2m_rand4 --scatter --steps 6 --warmup 2 --transports p2p --configs auto:auto:tma:-1:-1 8:0:tma:32:4
