#!/bin/bash
set -x
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "spmm or sage or gather or gcn" > $O/s3_pytest.log 2>&1; echo "pytest rc=$?"
timeout 200 python tools/kbench.py sage > $O/s3_sage_f32.jsonl 2> $O/s3_sage_f32.err
timeout 200 python tools/kbench.py sage --bf16 > $O/s3_sage_bf16.jsonl 2> $O/s3_sage_bf16.err
timeout 400 python tools/kbench.py spmm_reddit --Fs 602 128 16 --bf16 --reps 5 \
  --knobs '[{"spmm.rows_per_team":32},{"spmm.rows_per_team":8},{"spmm.rows_per_team":4},{"spmm.rows_per_team":2},{"spmm.rows_per_team":1},{"spmm.rows_per_team":0}]' \
  > $O/s3_spmm_reddit.jsonl 2> $O/s3_spmm_reddit.err
cat $O/s3_sage_f32.jsonl $O/s3_sage_bf16.jsonl | cut -c1-260
python - <<'PY'
import json
for l in open('gpurun_out/s3_spmm_reddit.jsonl'):
    d=json.loads(l)
    if 'ms' in d: print(d['F'], d['dtype'][6:], d['knobs'], d['rows_per_team'], round(d['ms'],2), round(d['gather_frac'],3))
PY
tail -3 $O/s3_pytest.log
