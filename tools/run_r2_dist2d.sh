#!/bin/bash
# 2-GPU: per-consumer timings of the scheduled step (which pass is below the 1-GPU SpMM rate?)
set -x
O=gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 \
  tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-1500; }
LOC="--p-local 0.8 --window 2000000 --scatter --steps 5 --warmup 2"
run 600 29543 r2e_full $LOC --consumers --phases --transports p2p --configs 4:0:tma:32:4 4:4:tma:32:4 1:1:tma:32:4
