#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r02d}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_robots.py tests/test_distance.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_quick.log
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2), 'launches', d['gpu_launches'])
"; tail -2 gpurun_out/err_$TAG.txt
done 2>&1 | tee gpurun_out/${TAG}_bench4.log
for w in wam_lab1_100k pr2_pr2test1_1m hironx_table_1m; do timeout 300 python tools/gpu_scale_n.py $w; done 2>&1 | tee gpurun_out/${TAG}_scale_n.log
