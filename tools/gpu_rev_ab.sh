#!/bin/bash
# A/B of the push order of the descent (CC_PUSH_REVERSED builds under build/): parity on the variant, then the workloads.
set -u
mkdir -p gpurun_out
TAG=${1:-rev}
CUDACOLLISION_LIB=$PWD/build/libcc_rev.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_robots.py tests/test_pair_tables.py tests/test_contacts.py -m gpu -x -q > gpurun_out/pytest_rev_$TAG.log 2>&1; echo "pytest (rev) rc=$?"; tail -3 gpurun_out/pytest_rev_$TAG.log
{
for lib in base rev revt8 revt32 base rev; do
  echo "=== $lib"
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
    CUDACOLLISION_LIB=$PWD/build/libcc_$lib.so timeout 300 python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_rev.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 3), 'M/s', round(d['ms_per_step'], 3), 'ms', 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2))
"; tail -1 gpurun_out/err_rev.txt
  done
done
} 2>&1 | tee gpurun_out/ab_rev_$TAG.log
