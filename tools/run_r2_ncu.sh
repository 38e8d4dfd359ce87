#!/bin/bash
# Round 2 ncu artefacts (1 GPU): bench launch list, full-set capture of the bench's gather launch and of the
# attention kernels on the full Reddit shape.  Reports are exported to CSV on the box (64 MiB cap on gpurun_out).
set -x
O=gpurun_out
T=/tmp/ncu; mkdir -p $T
BCMD="python bench.py --steps 3 --warmup 3 --skip-extra --no-cpu-baseline"
$BCMD > $O/r2n_bench_plain.json 2> $O/r2n_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_bench_launch_list.csv $BCMD > $O/r2n_ncu_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sage_tma_kernel -s 4 -c 2 -o $T/sage $BCMD > $O/r2n_ncu_b2.log 2>&1
GCMD="python tools/kbench.py gat --graph reddit_full --reps 1"
$GCMD > $O/r2n_gat_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gat_ -s 3 -c 4 -o $T/gat $GCMD > $O/r2n_ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gat_fwd -s 1 -c 1 -o $T/gat_bf16 $GCMD --bf16 > $O/r2n_ncu_g2.log 2>&1
for f in sage gat gat_bf16; do
  ncu -i $T/$f.ncu-rep --page raw --csv > $O/r02_prof_${f}_raw.csv 2>/dev/null
  ncu -i $T/$f.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_prof_${f}_source.csv.gz
  ls -la $T/$f.ncu-rep
done
du -sh $O
