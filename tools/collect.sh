#!/bin/bash
# Round collection on the GPU box: GPU tests, headline bench, reference arm, ncu launch list and one ncu --set full
# capture of the step's kernels.  usage: gpurun --timeout 1500 -- 'bash tools/collect.sh <tag> [quick]'
TAG=${1:-r1}
QUICK=${2:-}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi_$TAG.txt 2>&1
if [ -z "$QUICK" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $O/pytest_$TAG.log
  tail -3 $O/pytest_$TAG.log
fi
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -c 1500 $O/bench_$TAG.json
timeout 200 python tools/engine_timing.py > $O/engine_timing_$TAG.txt 2>&1; cat $O/engine_timing_$TAG.txt
if [ -z "$QUICK" ]; then
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference.json 2>> $O/bench_$TAG.err
  for w in s3dis_l1 s3dis_l5 modelnet_l2; do
    timeout 200 python bench.py --workload $w --steps 20 --warmup 5 > $O/bench_${TAG}_$w.json 2>> $O/bench_$TAG.err
  done
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_launch_$TAG.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on \
      -k regex:"k_gather_mma|k_backward_filter|k_neighbor_search|k_backward_lists|k_cloud_sort|k_group_items" -s 9 -c 9 -f \
      -o $O/prof_$TAG python tools/run_once.py 2 > $O/ncu_full_$TAG.log 2>&1
  tail -2 $O/ncu_full_$TAG.log
fi
