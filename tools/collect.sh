#!/bin/bash
# Round collection on the GPU box.  usage: gpurun --timeout 1500 -- 'bash tools/collect.sh <tag> [stage ...]'
# stages: test testall bench benchq ref layers nets ncu full tf32 f64 sanitize timing (default: test bench)
TAG=${1:-r2}
shift
STAGES=${@:-test bench}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi_$TAG.txt 2>&1
for S in $STAGES; do
case $S in
test)
  timeout 1200 python -m pytest tests -m gpu -q -x --timeout 300 > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $O/pytest_$TAG.log
  tail -15 $O/pytest_$TAG.log ;;
testall)
  timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $O/pytest_$TAG.log
  tail -40 $O/pytest_$TAG.log ;;
bench)
  timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench exit $?"; tail -c 3000 $O/bench_$TAG.json; tail -5 $O/bench_$TAG.err ;;
benchq)
  timeout 300 python bench.py --steps 20 --warmup 5 --no-sweep --no-cpu --no-parity > $O/benchq_$TAG.json 2> $O/benchq_$TAG.err; echo "bench exit $?"; tail -c 2500 $O/benchq_$TAG.json; tail -5 $O/benchq_$TAG.err ;;
ref)
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${TAG}_reference.json 2>> $O/bench_$TAG.err ;;
layers)
  for w in headline_b16 s3dis_l1 s3dis_l5 modelnet_l2; do
    timeout 200 python bench.py --workload $w --steps 20 --warmup 5 --no-parity > $O/bench_${TAG}_$w.json 2>> $O/bench_$TAG.err
  done ;;
nets)
  for w in seg_net cls_net; do
    timeout 300 python bench.py --workload $w --steps 20 --warmup 5 > $O/bench_${TAG}_$w.json 2>> $O/bench_$TAG.err; tail -c 1200 $O/bench_${TAG}_$w.json
  done ;;
timing)
  timeout 200 python tools/engine_timing.py > $O/engine_timing_$TAG.txt 2>&1; cat $O/engine_timing_$TAG.txt ;;
tf32)
  timeout 200 python tools/tf32_peak.py > $O/tf32_peak_$TAG.log 2>&1; cat $O/tf32_peak_$TAG.log ;;
ncu)
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-b16 --no-parity > $O/ncu_launch_$TAG.log 2>&1 ;;
full)
  timeout 500 ncu --set full --clock-control none --import-source on \
      -k regex:"k_gather_mma|k_backward_filter|k_neighbor_search|k_backward_lists|k_cloud_sort|k_group_items" -s 8 -c 8 -f \
      -o $O/prof_$TAG python tools/run_once.py 2 > $O/ncu_full_$TAG.log 2>&1
  tail -2 $O/ncu_full_$TAG.log ;;
f64)
  timeout 300 python tools/f64_timing.py > $O/f64_timing_$TAG.json 2> $O/f64_timing_$TAG.err; cat $O/f64_timing_$TAG.json ;;
sanitize)
  timeout 900 bash tools/sanitize.sh $TAG ;;
esac
done
