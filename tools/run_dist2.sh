#!/bin/bash
set -x
O=gpurun_out
for U in 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$U \
  tools/spmm_dist.py --nodes 8000000 --p-local 0.8 --window 200000 --scatter --transports p2p --check --cross-check --overlap-only \
  --dedicated 0 16 --halo-unroll $U > $O/s10_dist2_u$U.log 2>&1; echo rc=$?
grep -v "^\*\|OMP\|^W1\|^$" $O/s10_dist2_u$U.log | cut -c1-250
done
