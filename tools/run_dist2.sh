#!/bin/bash
set -x
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
  tools/spmm_dist.py --nodes 8000000 --p-local 0.8 --window 200000 --scatter --transports p2p --check --overlap-only \
  --dedicated 0 16 32 --halo-unroll 4 > $O/s5_dist2.log 2>&1; echo rc=$?
grep -v "^\*\|OMP\|^W1\|^$" $O/s5_dist2.log | cut -c1-330
