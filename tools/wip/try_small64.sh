#!/bin/bash
# Swaps the parked 64-channel version of small_channels.cu in ON THE GPU BOX, rebuilds, runs the GPU tests and the
# reference-model layer benches.  usage: gpurun -- bash tools/wip/try_small64.sh
cp tools/wip/small_channels_64ch.cu.wip pointwise_b200/csrc/small_channels.cu
python -m pointwise_b200.build --force > gpurun_out/wip_build.log 2>&1 || { tail -20 gpurun_out/wip_build.log; exit 1; }
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in s3dis_l5 s3dis_l1 modelnet_l2; do
  timeout 200 python bench.py --workload $w --steps 50 --warmup 10 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step']); [print('   ',k,v['avg_ms']) for k,v in list(d['kernels'].items())[:5]]"
done
