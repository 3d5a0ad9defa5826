#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-bvho}
run() {
  echo "=== $*"
  for w in ${WORKLOADS:-pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k}; do
    env "$@" timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_bvho.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 3), 'M/s', round(d['ms_per_step'], 3), 'ms', 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2))
"; tail -1 gpurun_out/err_bvho.txt
  done
}
{
run X=0
run CC_TUNE_BVH_ORDER=1
run CC_TUNE_BVH_ORDER=2
run CC_TUNE_BVH_ORDER=3
run CC_TUNE_BVH_COST=3
run X=0
timeout 300 python tools/gpu_probe.py 200000 pr2_pr2test1_1m
} 2>&1 | tee gpurun_out/ab_bvh_order_$TAG.log
