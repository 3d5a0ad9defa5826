#!/bin/bash
# ncu visit: launch list of the bench command + full captures of the two hot kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-prof}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_traverse -s 4 -c 1 -f -o gpurun_out/prof_trav_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_trav_$TAG.log 2>&1; echo "ncu trav rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_fk -s 4 -c 1 -f -o gpurun_out/prof_fk_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_fk_$TAG.log 2>&1; echo "ncu fk rc=$?"
grep -c cc_ gpurun_out/launches_$TAG.csv
