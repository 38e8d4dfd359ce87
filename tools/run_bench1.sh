#!/bin/bash
set -x
O=gpurun_out
SECONDS=0; timeout 900 python bench.py > $O/s7_bench.json 2> $O/s7_bench.err; echo "bench rc=$?"
echo "bench wall ${SECONDS}s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s7_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','wall_ms_per_step_incl_flush','gpu_launches','clocks')})
print('roofline',{k:d['roofline'][k] for k in ('achieved','frac','launch_ms','warm_l2')})
print('warm',d['warm_l2'])
print('e2e',{k:v for k,v in d['e2e'].items() if k not in('api','sync_note')})
print('e2e_ds',{k:v for k,v in d['e2e_device_sampling'].items() if k!='what'})
print('cpu',d.get('cpu_baseline'))
for k,v in d.get('other_configs',{}).items(): print(k,{a:b for a,b in v.items() if a!='config'})
for k,v in d.get('partitioned_spmm',{}).items(): print(k,{a:b for a,b in v.items() if a not in('graph','transport')})
PY
