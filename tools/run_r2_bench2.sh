#!/bin/bash
# round 2: the driver's N=2 launch of bench.py (both arms), after a quick HAN re-check
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "han or gather_reduce_backward_bf16" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2_bench_2gpu.json; tail -5 gpurun_out/r2_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_bench_2gpu_ref.json 2> gpurun_out/r2_bench_2gpu_ref.err; echo "ref rc=$?"
tail -c 800 gpurun_out/r2_bench_2gpu_ref.json
