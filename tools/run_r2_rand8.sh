#!/bin/bash
# Round 2, 8 GPUs, random graph fp32: which of (two-pass count, mover CTAs) the exchange-bound case wants
O=gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 \
  tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-330; }
run 400 29571 r2i_rand --scatter --steps 6 --warmup 2 --transports p2p \
  --configs auto:auto:tma:-1:-1 8:0:tma:32:4:0:fused:0 8:8:tma:48:4:0:fused:0 8:0:tma:48:4:0:fused:0 8:4:tma:32:4:0:fused:0
