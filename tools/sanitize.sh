#!/bin/bash
# compute-sanitizer over a reduced subset of the GPU tests: memcheck on everything in the subset,
# racecheck on the shared-memory heavy kernels.  usage (on the GPU box): bash tools/sanitize.sh <tag>
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
SUBSET='tests/test_gpu_parity.py::test_known_answers tests/test_gpu_parity.py::test_ragged_cloud_sizes_through_tiles tests/test_gpu_tc.py'
K='test_tensor_core_engine_matches_oracle_and_simt and (64-128 or 64-64 or 256-256 or 32-32)'
compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python tools/sanitize_run.py > $O/sanitize_memcheck_$TAG.log 2>&1
echo "memcheck exit $?" >> $O/sanitize_memcheck_$TAG.log
tail -8 $O/sanitize_memcheck_$TAG.log
compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 20 python tools/sanitize_run.py small > $O/sanitize_racecheck_$TAG.log 2>&1
echo "racecheck exit $?" >> $O/sanitize_racecheck_$TAG.log
tail -8 $O/sanitize_racecheck_$TAG.log
