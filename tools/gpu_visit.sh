#!/bin/bash
# Full GPU-box visit: parity tests, the three single-GPU workloads, the reference arm, the ncu launch list of the bench
# command and full captures of the two hot kernels.  Outputs land in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
TAG=${1:-r01b}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv 2>&1
nproc > gpurun_out/nproc_$TAG.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/nproc_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_gpu_$TAG.log | tail -6
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"
  cut -c1-300 gpurun_out/bench_${w}_$TAG.json; tail -2 gpurun_out/bench_${w}_$TAG.err
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_traverse -s 12 -c 1 -f -o gpurun_out/prof_trav_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_trav_$TAG.log 2>&1; echo "ncu trav rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_fk -s 12 -c 1 -f -o gpurun_out/prof_fk_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_fk_$TAG.log 2>&1; echo "ncu fk rc=$?"
ls -la gpurun_out | head -40
