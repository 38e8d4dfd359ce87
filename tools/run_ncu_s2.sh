#!/bin/bash
# ncu artefacts for round 1, session 2 (each command first exits 0 without ncu)
set -x
O=gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --skip-extra --no-cpu-baseline > $O/s8_plain.json 2> $O/s8_plain.err || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/s8_launches.csv \
  python bench.py --steps 5 --warmup 3 --skip-extra --no-cpu-baseline > $O/s8_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sage_tma --launch-skip 6 -c 2 -o $O/s8_prof_sage -f \
  python bench.py --steps 5 --warmup 3 --skip-extra --no-cpu-baseline > $O/s8_ncu2.log 2>&1
timeout 200 python tools/kbench.py spmm_reddit --Fs 602 --reps 2 > $O/s8_kb.jsonl 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmm_rbs --launch-skip 3 -c 1 -o $O/s8_prof_spmm_reddit -f \
  python tools/kbench.py spmm_reddit --Fs 602 --reps 2 > $O/s8_ncu3.log 2>&1
ls -la $O/s8_*
