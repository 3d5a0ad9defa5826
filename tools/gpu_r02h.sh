#!/bin/bash
# Round-2 visit h: work sharing of heavy configurations.  Every step under a short timeout (a hang must not become a strike).
set -u
mkdir -p gpurun_out
TAG=${1:-r02h}
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick.log 2>&1; rc=$?; echo "quick pytest rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_quick.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 240 python tools/gpu_scale_n.py pr2_pr2test1_1m 12500,100000,400000,1000000 2>&1 | tee gpurun_out/${TAG}_scale_first.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
for knobs in "CC_TUNE_SHARE=0" "CC_TUNE_SHARE=256" "CC_TUNE_SHARE=0" "CC_TUNE_SHARE=256"; do
  echo "=== $knobs"
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
    env $knobs timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2), 'dynamic', round(d['config']['dynamic_scene_configs_per_s']/1e6, 2))
"; tail -2 gpurun_out/err_$TAG.txt
  done
done 2>&1 | tee gpurun_out/${TAG}_ab_share.log
for b in 0 256; do for w in wam_lab1_100k pr2_pr2test1_1m; do CC_TUNE_SHARE=$b timeout 300 python tools/gpu_scale_n.py $w 1000,4000,12500,25000,50000,100000,200000,400000,1000000; done; done 2>&1 | tee gpurun_out/${TAG}_scale_n.log
for b in 0 256; do CC_TUNE_SHARE=$b timeout 300 python bench.py --workload wam_wamtest1_birrt100 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200; done
