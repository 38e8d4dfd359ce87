#!/bin/bash
# Round 2, 8-GPU: interleaved dealing of the wave mover's segments (all peers fed at once) vs rotated peers
set -x
O=gpurun_out
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 \
  tools/spmm_dist.py "${@:4}" > $O/$3.log 2>&1; echo "$3 rc=$?"; grep -v "^\*\|OMP\|^W1\|^$" $O/$3.log | cut -c1-400; }
run 420 29551 r2g_loc --p-local 0.8 --window 2000000 --scatter --steps 8 --warmup 2 --phases --transports p2p \
  --configs 4:4:tma:48:4:0:fused:0 4:4:tma:48:4:0:fused:16 4:4:tma:48:4:0:fused:4 4:4:tma:32:4:0:fused:16 4:4:tma:64:4:0:fused:16 4:4:tma:48:4:0:fused:64 1:1:tma:48:4:0:fused:16 2:2:tma:48:4:0:fused:16
run 300 29552 r2g_rand --scatter --steps 6 --warmup 2 --phases --transports p2p \
  --configs 8:0:tma:32:4:0:fused:0 8:0:tma:32:4:0:fused:16 8:0:tma:48:4:0:fused:16 8:0:tma:24:4:0:fused:16
