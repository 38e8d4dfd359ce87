#!/bin/bash
# Round 2, 1-GPU call: whole GPU suite (new: fused train epilogue + seeded dropout, goldens at config size, reference
# substitution, property tests), GAT micro-benchmarks, a short bench line.
set -x
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 $O/r2i_pytest.log | cut -c1-250
timeout 600 python tools/kbench.py gat --graph reddit_full --reps 5 > $O/r2i_gat_full.jsonl 2> $O/r2i_gat.err; cat $O/r2i_gat_full.jsonl
timeout 600 python tools/kbench.py gat --graph reddit_full --reps 5 --bf16 > $O/r2i_gat_full_bf16.jsonl 2>> $O/r2i_gat.err; cat $O/r2i_gat_full_bf16.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2i_bench.json 2> $O/r2i_bench.err; echo "bench rc=$?"; cut -c1-700 $O/r2i_bench.json; tail -3 $O/r2i_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 --skip-extra > $O/r2i_bench_ref.json 2> $O/r2i_bench_ref.err; echo "ref rc=$?"; cut -c1-900 $O/r2i_bench_ref.json
