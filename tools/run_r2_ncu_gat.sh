#!/bin/bash
# ncu: per-launch durations of the GAT kernels on the full Reddit shape, then one full-set capture of the backward pair
set -x
O=gpurun_out
CMD="python tools/kbench.py gat --graph reddit_full --reps 1"
$CMD > $O/r2n_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gat_ -c 40 --csv --log-file $O/r2n_gat_launches.csv $CMD > $O/r2n_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gat_bwd -s 4 -c 3 -o $O/r2n_prof_gat_bwd $CMD > $O/r2n_ncu2.log 2>&1
CMDB="python tools/kbench.py gat --graph reddit_full --reps 1 --bf16"
ncu --set full --clock-control none --import-source on -k regex:gat_fwd -s 2 -c 2 -o $O/r2n_prof_gat_fwd_bf16 $CMDB > $O/r2n_ncu3.log 2>&1
ls -la $O/r2n_*
