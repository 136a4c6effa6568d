#!/bin/bash
# Times the production library and every variant under pointwise_b200/lib/variants/ in one GPU call.
# usage: bash tools/ab_variants.sh [workload] [engine flags ...]
timeout 120 python tools/ab_backward.py "$@" 2>&1 | grep -v Warning
for v in pointwise_b200/lib/variants/*.so; do
  CONV3P_LIB=$v timeout 120 python tools/ab_backward.py "$@" 2>&1 | grep -v Warning
done
