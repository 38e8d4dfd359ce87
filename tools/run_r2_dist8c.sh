#!/bin/bash
# Round 2, 8-GPU sweep #2: single-launch wave mover (in-kernel flags), wider mover geometries (the exchange, not the
# local pass, is the critical path at 8 GPUs), per-consumer timings; locality + random graph.
set -x
O=gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 \
  tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-500; }
run 420 29551 r2f_loc --p-local 0.8 --window 2000000 --scatter --steps 8 --warmup 2 --phases --consumers --transports p2p \
  --configs 4:4:tma:48:4 4:4:tma:64:4 4:4:tma:96:4 4:4:tma:148:4 4:4:tma:148:2 1:1:tma:64:4 1:1:tma:148:4 2:2:tma:64:4 8:8:tma:64:4 4:3:tma:64:4 4:4:tma:48:4:0:sep
run 420 29552 r2f_rand --scatter --steps 6 --warmup 2 --phases --transports p2p \
  --configs 8:0:tma:32:4 8:0:tma:64:4 8:0:tma:148:4 16:0:tma:64:4 4:0:tma:64:4 2:0:tma:64:4
