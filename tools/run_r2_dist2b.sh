#!/bin/bash
# Round 2, 2-GPU call: GPU unit tests, then the wave-pipelined partitioned SpMM — correctness on a small graph
# (every row against a 1-GPU SpMM, forward + backward, all transports) and the schedule sweep at full size.
set -x
O=gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r2b_pytest.log
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 \
  tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-700; }
SMALL="--nodes 8000000 --p-local 0.8 --window 200000 --scatter --full-check --backward"
run 300 29541 r2b_small $SMALL --transports p2p ce nccl --configs 1:1:tma:0:1 4:auto:tma:0:1 4:0:vector:0:0 3:2:tma:16:4
run 200 29542 r2b_small_bf16 $SMALL --dtype bf16 --transports p2p --configs 4:auto:tma:0:1
LOC="--p-local 0.8 --window 2000000 --scatter --steps 5 --warmup 2"
run 600 29543 r2b_full $LOC --phases --transports p2p --configs 1:1:tma:0:1 1:1:tma:0:2 1:1:tma:32:4 1:1:vector:0:0:32 1:0:tma:0:1 4:auto:tma:0:1 4:1:tma:0:1 4:2:tma:0:1 4:0:tma:0:1 8:auto:tma:0:1
