#!/bin/bash
# 1-GPU validation: full GPU parity suite, smoke, kernel micro-benches touched this session
set -x
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/s4_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $O/s4_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/s4_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/s4_smoke.log
timeout 200 python tools/kbench.py sage --knobs '[{"sage.packed_add":0},{"sage.packed_add":1}]' > $O/s4_sage_f32.jsonl 2> $O/s4_sage_f32.err
timeout 200 python tools/kbench.py sage --bf16 --knobs '[{"sage.packed_add":0},{"sage.packed_add":1}]' > $O/s4_sage_bf16.jsonl 2> $O/s4_sage_bf16.err
cat $O/s4_sage_f32.jsonl $O/s4_sage_bf16.jsonl | cut -c1-250
