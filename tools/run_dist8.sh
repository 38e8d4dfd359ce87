#!/bin/bash
set -x
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 \
  tools/spmm_dist.py --p-local 0.8 --window 2000000 --scatter --transports p2p --overlap-only --cross-check \
  --dedicated 0 16 24 32 48 --halo-unroll 4 8 --steps 6 --warmup 2 > $O/s11_dist8.log 2>&1; echo rc=$?
grep -v "^\*\|OMP\|^W1\|^$" $O/s11_dist8.log | cut -c1-255
