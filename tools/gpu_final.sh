#!/bin/bash
# Closing visit at HEAD: GPU parity tests, smoke, every bench workload (+ reference arms), ncu launch list and full captures.
set -u
mkdir -p gpurun_out
TAG=${1:-r01D}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv 2>&1
nproc > gpurun_out/nproc_$TAG.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/nproc_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k wam_wamtest1_birrt100 wam_lab1_distance_100k; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"
  python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${w}_$TAG.json').read().strip().splitlines()[-1]); print('  ', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'cpu', d.get('cpu_baseline',{}).get('value'))"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; tail -c 300 gpurun_out/bench_ref_$TAG.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_traverse -s 10 -c 1 -f -o gpurun_out/prof_trav_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_trav_$TAG.log 2>&1; echo "ncu trav rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_fk -s 40 -c 1 -f -o gpurun_out/prof_fk_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_fk_$TAG.log 2>&1; echo "ncu fk rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_traverse -s 10 -c 1 -f -o gpurun_out/prof_trav_clutter_$TAG \
    python bench.py --workload wam_clutter200_1250k --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_trav_clutter_$TAG.log 2>&1; echo "ncu trav clutter rc=$?"
ls -la gpurun_out | tail -8
