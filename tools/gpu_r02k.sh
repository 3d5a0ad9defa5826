#!/bin/bash
# Round-2 visit k: broadphase as its own kernel (cc_broad_kernel: one bit per pair and configuration) against the broadphase inside
# the traversal warp; then a smaller LIFO (224 items: 7 blocks/SM fit a 196 KB carve-out, L1 grows from 28 to 60 KB).
set -u
mkdir -p gpurun_out
TAG=${1:-r02k}
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_pair_tables.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick.log 2>&1; rc=$?; echo "quick pytest rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_quick.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/${TAG}_pytest_quick.log; exit 1; fi
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
for knobs in "CC_TUNE_CAND=0" "CC_TUNE_CAND=1" "CC_TUNE_CAND=0" "CC_TUNE_CAND=1"; do echo "=== $knobs"; bench4 $knobs; done 2>&1 | tee gpurun_out/${TAG}_ab_cand.log
for knobs in "CUDACOLLISION_LIB=$PWD/build/libcc_stack224.so" "CUDACOLLISION_LIB=$PWD/build/libcc_stack224.so CC_TUNE_CARVEOUT=86" "X=0"; do echo "=== $knobs"; bench4 $knobs; done 2>&1 | tee gpurun_out/${TAG}_ab_stack224.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python tools/gpu_scale_n.py pr2_pr2test1_1m 100000,400000,1000000 2>&1 | tee gpurun_out/${TAG}_scale_n.log
