#!/bin/bash
# Matrix of BVH build knobs (host-side only; verdicts do not depend on them) on the three robot workloads.
set -u
mkdir -p gpurun_out
TAG=${1:-bvhk}
run() {
  echo "=== $*"
  for w in ${WORKLOADS:-pr2_pr2test1_1m wam_lab1_100k hironx_table_1m}; do
    env "$@" timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_bvhk.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 3), 'M/s', round(d['ms_per_step'], 3), 'ms')
"; tail -1 gpurun_out/err_bvhk.txt
  done
}
{
run X=0
run CC_TUNE_BVH_BINS=32
run CC_TUNE_BVH_BINS=64
run CC_TUNE_BVH_BINS=8
run CC_TUNE_BVH_COST=1
run CC_TUNE_BVH_COST=2
run CC_TUNE_BVH_COST=3
run CC_TUNE_BVH_SIZE=1
run CC_TUNE_BVH_SIZE=2
run CC_TUNE_BVH_SIZE=3
run X=0
} 2>&1 | tee gpurun_out/ab_bvh_knobs_$TAG.log
