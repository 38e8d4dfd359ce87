#!/bin/bash
# ncu: per-launch durations of the GAT kernels on the full Reddit shape, then one full-set capture of the backward pair.
# Reports are exported to CSV on the box (raw + source pages); the .ncu-rep is kept only when small (64 MiB cap).
set -x
O=gpurun_out
T=/tmp/ncu; mkdir -p $T
CMD="python tools/kbench.py gat --graph reddit_full --reps 1"
$CMD > $O/r2n_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gat_ -c 40 --csv --log-file $O/r2n_gat_launches.csv $CMD > $O/r2n_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gat_bwd -s 4 -c 3 -o $T/gat_bwd $CMD > $O/r2n_ncu2.log 2>&1
CMDB="python tools/kbench.py gat --graph reddit_full --reps 1 --bf16"
ncu --set full --clock-control none --import-source on -k regex:gat_fwd -s 2 -c 1 -o $T/gat_fwd_bf16 $CMDB > $O/r2n_ncu3.log 2>&1
for f in gat_bwd gat_fwd_bf16; do
  ncu -i $T/$f.ncu-rep --page raw --csv > $O/r2n_prof_${f}_raw.csv 2>/dev/null
  ncu -i $T/$f.ncu-rep --page source --csv > $O/r2n_prof_${f}_source.csv 2>/dev/null
  ls -la $T/$f.ncu-rep
done
gzip -f $O/r2n_prof_*_source.csv
du -sh $O; ls -la $O | tail
