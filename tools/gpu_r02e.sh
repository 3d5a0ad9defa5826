#!/bin/bash
# Round-2 visit e: heavy-configuration deferral (iteration budget + split relaunch).  Parity with a tiny budget (almost everything
# that collides deeply is deferred), parity at the default, then budget A/B on batch-size scaling and the bench workloads.
set -u
mkdir -p gpurun_out
TAG=${1:-r02e}
CC_TUNE_DEFER=6 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu_defer6.log 2>&1; echo "pytest (budget 6) rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu_defer6.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
for b in 0 24 48 64 96; do
  echo "=== CC_TUNE_DEFER=$b"
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
    CC_TUNE_DEFER=$b timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2), 'launches', d['gpu_launches'])
"; tail -2 gpurun_out/err_$TAG.txt
  done
done 2>&1 | tee gpurun_out/${TAG}_ab_defer.log
for b in 0 64; do for w in wam_lab1_100k pr2_pr2test1_1m; do CC_TUNE_DEFER=$b timeout 300 python tools/gpu_scale_n.py $w 1000,4000,12500,25000,50000,100000,200000,400000,1000000; done; done 2>&1 | tee gpurun_out/${TAG}_scale_n.log
timeout 300 python bench.py --workload wam_wamtest1_birrt100 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-300
CC_TUNE_DEFER=0 timeout 300 python bench.py --workload wam_wamtest1_birrt100 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-300
