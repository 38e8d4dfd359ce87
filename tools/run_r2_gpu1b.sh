#!/bin/bash
# Round 2, 1-GPU call: GPU suite (factorised GAT backward, GTN, reference substitution), GAT micro-benchmarks
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/r2h_pytest.log | cut -c1-250
for g in reddit_d100 reddit_full; do
timeout 600 python tools/kbench.py gat --graph $g --reps 5 > $O/r2h_gat_$g.jsonl 2> $O/r2h_gat.err; cat $O/r2h_gat_$g.jsonl
timeout 600 python tools/kbench.py gat --graph $g --reps 5 --bf16 > $O/r2h_gat_${g}_bf16.jsonl 2>> $O/r2h_gat.err; cat $O/r2h_gat_${g}_bf16.jsonl
done
timeout 300 python tools/kbench.py gat --graph reddit_full --reps 3 --knobs '[{"gat.bwd_stage_edges":32},{"gat.bwd_stage_edges":128}]' > $O/r2h_gat_se.jsonl 2>> $O/r2h_gat.err; cat $O/r2h_gat_se.jsonl
timeout 300 python tools/kbench.py gat --graph acm_dense --reps 5 > $O/r2h_gat_acm.jsonl 2>> $O/r2h_gat.err; cat $O/r2h_gat_acm.jsonl
tail -3 $O/r2h_gat.err
