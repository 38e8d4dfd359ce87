#!/bin/bash
# Round 2, 8-GPU sweep (~70 s of box time = ~10 GPU-minutes): papers100M-shaped locality graph,
# vector-store push (reference) vs TMA mover on 8..32 dedicated SMs, then the copy-engine transport.
#   gpurun --gpus 8 --timeout 500 -- 'bash tools/run_r2_dist8.sh'
set -x
O=gpurun_out
COMMON="--p-local 0.8 --window 2000000 --scatter --overlap-only --cross-check --halo-unroll 4 --steps 6 --warmup 2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 \
  tools/spmm_dist.py $COMMON --transports p2p --tma --dedicated 8 16 24 32 > $O/r2_d8_tma.log 2>&1; echo rc=$?
grep -v "^\*\|OMP\|^W1\|^$" $O/r2_d8_tma.log | cut -c1-260
