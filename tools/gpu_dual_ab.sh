#!/bin/bash
# A/B of the dual-descent items (CC_DUAL_RATIO builds under build/): parity on one variant, then the workloads on all.
set -u
mkdir -p gpurun_out
TAG=${1:-dual}
CUDACOLLISION_LIB=$PWD/build/libcc_dual8.so timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_robots.py tests/test_gpu_parity_clutter.py tests/test_pair_tables.py tests/test_contacts.py tests/test_grabbed.py -m gpu -x -q > gpurun_out/pytest_dual_$TAG.log 2>&1; echo "pytest (dual8) rc=$?"; tail -3 gpurun_out/pytest_dual_$TAG.log
{
for lib in base dual8 dual4 dual64 dualinf dual8b6 base; do
  echo "=== $lib"
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
    CUDACOLLISION_LIB=$PWD/build/libcc_$lib.so timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_dual.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 3), 'M/s', round(d['ms_per_step'], 3), 'ms', 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2))
"; tail -1 gpurun_out/err_dual.txt
  done
done
} 2>&1 | tee gpurun_out/ab_dual_$TAG.log
