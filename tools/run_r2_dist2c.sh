#!/bin/bash
# Round 2, 2-GPU call: single-launch wave mover (in-kernel arrival flags) — correctness on the small graph (every row
# vs a 1-GPU SpMM, forward + backward), then fused vs per-wave launches and mover geometries at full size.
set -x
O=gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 \
  tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-420; }
SMALL="--nodes 8000000 --p-local 0.8 --window 200000 --scatter --full-check --backward"
run 300 29541 r2d_small $SMALL --transports p2p --configs 4:auto:tma:0:1 4:2:tma:32:4 8:0:tma:16:4 1:1:tma:0:2 4:auto:tma:0:1:0:sep
run 200 29542 r2d_small_bf16 $SMALL --dtype bf16 --transports p2p --configs 4:auto:tma:32:4
LOC="--p-local 0.8 --window 2000000 --scatter --steps 5 --warmup 2"
run 600 29543 r2d_full $LOC --phases --transports p2p --configs 4:0:tma:0:1 4:0:tma:0:1:0:sep 4:0:tma:32:4 4:0:tma:48:4 4:0:tma:64:2 4:0:tma:74:2 8:0:tma:32:4 2:0:tma:32:4
