#!/bin/bash
set -x
O=gpurun_out
timeout 300 python tools/kbench.py spmm_reddit --Fs 602 256 --bf16 --reps 5 --knobs '[{"spmm.ctas3":0},{"spmm.ctas3":1}]' > $O/s9_spmm_ctas3.jsonl 2> $O/s9_spmm_ctas3.err
timeout 300 python tools/kbench.py gat --graph reddit_full --reps 3 > $O/s9_gat_full.jsonl 2> $O/s9_gat_full.err
timeout 400 python tools/kbench.py spmm_papers --Fs 128 --only-bf16 --reps 3 > $O/s9_papers_bf16.jsonl 2> $O/s9_papers_bf16.err
python - <<'PY'
import json
for f in ('s9_spmm_ctas3','s9_gat_full','s9_papers_bf16'):
    for l in open(f'gpurun_out/{f}.jsonl'):
        d=json.loads(l)
        print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k not in ('planned','compulsory_gbs','X_mb','ms_best')})
PY
tail -2 $O/s9_gat_full.err $O/s9_papers_bf16.err
