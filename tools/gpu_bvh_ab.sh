#!/bin/bash
# A/B of the BVH split rule (CC_TUNE_BVH=mid vs the default binned SAH): parity first, then the big workloads.
set -u
mkdir -p gpurun_out
TAG=${1:-bvh}
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_robots.py tests/test_pair_tables.py tests/test_distance.py tests/test_contacts.py -m gpu -x -q > gpurun_out/pytest_bvh_$TAG.log 2>&1; echo "pytest (sah) rc=$?"; tail -3 gpurun_out/pytest_bvh_$TAG.log
for mode in mid sah mid sah; do
  echo "=== CC_TUNE_BVH=$mode"
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k wam_lab1_distance_100k; do
    CC_TUNE_BVH=$mode timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_bvh.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 3), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 3))
"; tail -1 gpurun_out/err_bvh.txt
  done
done 2>&1 | tee gpurun_out/ab_bvh_$TAG.log
