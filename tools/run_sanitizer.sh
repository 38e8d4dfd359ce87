#!/bin/bash
# compute-sanitizer over the smoke + a bounded subset of the GPU parity suite (SURVEY.md §5 "race detection").
# ONE tool per gpurun call (B200_PROFILING.md: several sanitizer tools in one call left a GPU unusable):
#     gpurun --timeout 1500 -- 'bash tools/run_sanitizer.sh memcheck'
#     gpurun --timeout 1500 -- 'bash tools/run_sanitizer.sh racecheck'
# The plain run goes first; the sanitizer only runs if it exits 0.  Output: gpurun_out/sanitizer_<tool>.log
# (copy the summary lines into profiles/ when quoting them).
# NOTE (r02): compute-sanitizer is CLOSED on this pool (the wrapper refuses it: runs under it have left GPUs needing a
# reset) — the script is kept for a pool where it is open.  In-tree substitutes: ABI-side index validation, the
# out-of-range / ragged / empty-row GPU tests and tests/test_gpu_properties.py.
set -x
TOOL=${1:-memcheck}
O=gpurun_out
SUBSET=(tests/test_gpu_partition.py tests/test_gpu_parity.py -k "(spmm_f32_shapes and 128) or (gather_reduce_f32 and 602-10 and mean) or gat_small or halo_push or peer_flags or spmm_ex_row_subset or out_of_range")
timeout 600 python __graft_entry__.py --smoke > $O/sanitizer_plain.log 2>&1 && \
timeout 600 python -m pytest -m gpu -x -q "${SUBSET[@]}" >> $O/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -20 $O/sanitizer_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 9 --launch-timeout 0 python __graft_entry__.py --smoke > $O/sanitizer_${TOOL}.log 2>&1; echo "smoke under $TOOL rc=$?"
timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 9 python -m pytest -m gpu -x -q "${SUBSET[@]}" >> $O/sanitizer_${TOOL}.log 2>&1; echo "pytest subset under $TOOL rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Invalid|Race" $O/sanitizer_${TOOL}.log | tail -20
