#!/bin/bash
# Round 2, 1-GPU call: full GPU parity suite, smoke, GAT micro-benchmarks (transposed-order stash, bf16), bench line.
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r2g_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/r2g_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r2g_smoke.log
timeout 600 python tools/kbench.py gat --graph reddit_d100 --reps 5 > $O/r2g_gat_d100.jsonl 2> $O/r2g_gat.err; cat $O/r2g_gat_d100.jsonl
timeout 600 python tools/kbench.py gat --graph reddit_full --reps 5 > $O/r2g_gat_full.jsonl 2>> $O/r2g_gat.err; cat $O/r2g_gat_full.jsonl
timeout 600 python tools/kbench.py gat --graph reddit_full --reps 5 --bf16 > $O/r2g_gat_full_bf16.jsonl 2>> $O/r2g_gat.err; cat $O/r2g_gat_full_bf16.jsonl
timeout 900 python bench.py --steps 100 --warmup 5 > $O/r2g_bench.json 2> $O/r2g_bench.err; echo "bench rc=$?"; cut -c1-1500 $O/r2g_bench.json; tail -5 $O/r2g_bench.err
