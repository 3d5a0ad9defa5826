#!/bin/bash
# Round-2 visit A: descent iterations with <= 16 items give every item two lanes (one box test per lane) vs one lane per item.
set -u
mkdir -p gpurun_out
TAG=${1:-r02A}
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_pair_tables.py tests/test_contacts.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick.log 2>&1; rc=$?; echo "quick pytest rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_quick.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/${TAG}_pytest_quick.log; exit 1; fi
bench4() {
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
    env "$@" timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2), 'dynamic', round(d['config']['dynamic_scene_configs_per_s']/1e6, 2), 'launches', d['gpu_launches'])
"; tail -2 gpurun_out/err_$TAG.txt
  done
}
for knobs in "X=twolanes" "CUDACOLLISION_LIB=$PWD/build/libcc_nosplitlanes.so" "X=twolanes" "CUDACOLLISION_LIB=$PWD/build/libcc_nosplitlanes.so"; do echo "=== $knobs"; bench4 $knobs; done 2>&1 | tee gpurun_out/${TAG}_ab_split_lanes.log
timeout 300 python bench.py --workload wam_wamtest1_birrt100 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-160
