#!/bin/bash
# Round 2, 1-GPU call: GPU suite after the split-GEMM / GTN / property fixes, short bench (e2e with the tensor-core product)
set -x
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r2j_pytest.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2j_bench.json 2> $O/r2j_bench.err; echo "bench rc=$?"; cut -c1-300 $O/r2j_bench.json; tail -3 $O/r2j_bench.err
