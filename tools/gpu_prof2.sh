#!/bin/bash
# ncu captures of the distance kernel and of the split traversal (small batch), launch list of the BiRRT bench.
set -u
mkdir -p gpurun_out
TAG=${1:-r01k}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_distance -s 3 -c 1 -f -o gpurun_out/prof_dist_$TAG \
    python bench.py --workload wam_lab1_distance_100k --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dist_$TAG.log 2>&1; echo "ncu dist rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_birrt_$TAG.csv \
    python bench.py --workload wam_wamtest1_birrt100 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_birrt_$TAG.log 2>&1; echo "ncu birrt rc=$?"
grep -c cc_ gpurun_out/launches_birrt_$TAG.csv
ls -la gpurun_out | tail -5
