#!/bin/bash
# Round 2, 8-GPU sweep: wave-pipelined partitioned SpMM on the papers100M-shaped graph (locality + random),
# schedule (waves : two-pass chunks) x mover geometry; every line carries the float64 spot check.
set -x
O=gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 \
  tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-600; }
run 420 29551 r2c_loc --p-local 0.8 --window 2000000 --scatter --steps 8 --warmup 2 --phases --backward --transports p2p \
  --configs 4:4:tma:32:4 4:4:tma:48:4 4:4:tma:0:1 4:3:tma:32:4 4:2:tma:32:4 4:auto:tma:32:4 1:1:tma:32:4 1:1:vector:0:0:32 8:auto:tma:32:4
run 420 29552 r2c_rand --scatter --steps 6 --warmup 2 --phases --transports p2p \
  --configs 4:auto:tma:32:4 4:0:tma:32:4 8:0:tma:32:4 8:0:tma:64:2 4:0:tma:0:1 1:1:vector:0:0:32
