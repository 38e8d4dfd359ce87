#!/bin/bash
# round 2, final 1-GPU pass: full GPU suite, default bench (both arms), then compute-sanitizer memcheck on smoke()
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r2_final_pytest.log
timeout 500 python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_final_bench_ref.json 2> $O/r2_final_bench_ref.err; echo "ref rc=$?"
timeout 200 python __graft_entry__.py --smoke > $O/sanitizer_plain.log 2>&1; echo "smoke rc=$?"
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|smoke" $O/sanitizer_memcheck.log | tail -5
