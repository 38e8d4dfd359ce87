#!/bin/bash
# Round 2 probe (2 GPUs): is the TMA mover correct, and how many SMs does each transport need for NVLink rate?
set -x
O=gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 \
  tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-300; }
run 200 29541 r2p_small --nodes 8000000 --p-local 0.8 --window 200000 --scatter --check --cross-check --overlap-only --halo-unroll 4 --transports p2p --tma --dedicated 0 8 16
run 200 29542 r2p_ce --nodes 8000000 --p-local 0.8 --window 200000 --scatter --check --cross-check --overlap-only --halo-unroll 4 --transports ce
run 400 29543 r2p_full --p-local 0.8 --window 2000000 --scatter --cross-check --overlap-only --halo-unroll 4 --transports p2p --tma --dedicated 4 8 16 24 32 --phases --steps 4 --warmup 2
