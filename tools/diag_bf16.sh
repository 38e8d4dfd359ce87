#!/bin/bash
# bf16 diagnostics (round 1, session 2): new SAGE id-prefetch producer + where the bf16 SpMM time goes
set -x
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "sage or gather" > $O/s2_pytest_sage.log 2>&1; echo "pytest rc=$?"
timeout 200 python tools/kbench.py sage > $O/s2_sage_f32.jsonl 2> $O/s2_sage_f32.err
timeout 200 python tools/kbench.py sage --bf16 > $O/s2_sage_bf16.jsonl 2> $O/s2_sage_bf16.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s2_spmm_launches.csv \
  python tools/kbench.py spmm_reddit --Fs 602 --bf16 --reps 2 > $O/s2_spmm_launches.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm --launch-skip 9 -c 3 -o $O/s2_prof_spmm_bf16 -f \
  python tools/kbench.py spmm_reddit --Fs 602 --only-bf16 --reps 2 > $O/s2_prof_spmm_bf16.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sage_tma --launch-skip 3 -c 1 -o $O/s2_prof_sage_bf16 -f \
  python tools/kbench.py sage --bf16 --reps 2 > $O/s2_prof_sage_bf16.out 2>&1
cat $O/s2_sage_f32.jsonl $O/s2_sage_bf16.jsonl | cut -c1-300
tail -3 $O/s2_pytest_sage.log
