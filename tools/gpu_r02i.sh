#!/bin/bash
# Round-2 visit i: analytic primitives tests, full GPU suite at HEAD, the four collision workloads, ncu launch list + full captures.
set -u
mkdir -p gpurun_out
TAG=${1:-r02i}
timeout 600 python -m pytest tests/test_analytic_primitives.py -m gpu -x -q > gpurun_out/${TAG}_pytest_analytic.log 2>&1; echo "analytic pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest_analytic.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2), 'dynamic', round(d['config']['dynamic_scene_configs_per_s']/1e6, 2))
"; tail -2 gpurun_out/err_$TAG.txt
done 2>&1 | tee gpurun_out/${TAG}_bench4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_traverse -s 6 -c 1 -f -o gpurun_out/${TAG}_prof_trav \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_trav.log 2>&1; echo "ncu trav rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_fk -s 14 -c 1 -f -o gpurun_out/${TAG}_prof_fk \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_fk.log 2>&1; echo "ncu fk rc=$?"
timeout 300 python tools/gpu_timeline.py pr2_pr2test1_1m > gpurun_out/${TAG}_timeline_pr2.json 2> gpurun_out/err_tl.txt; tail -2 gpurun_out/err_tl.txt
