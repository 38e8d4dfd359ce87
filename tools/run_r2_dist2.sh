#!/bin/bash
# Round 2, first multi-GPU call (2 GPUs, ~1 GPU-minute x2): correctness of the two transports that were written
# after round 1's GPU budget was spent -- the TMA mover (halo.tma) and the copy-engine transport ("ce").
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/run_r2_dist2.sh'
set -x
O=gpurun_out
COMMON="--nodes 8000000 --p-local 0.8 --window 200000 --scatter --check --cross-check --overlap-only --halo-unroll 4"
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 \
  tools/spmm_dist.py $COMMON "${@:3}" > $O/$2.log 2>&1; echo "$2 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$2.log | cut -c1-260; }
run 29541 r2_d2_tma  --transports p2p --tma --dedicated 0 8 16
run 29542 r2_d2_ce   --transports ce
