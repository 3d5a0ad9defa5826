#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-c5}
timeout 600 python -m pytest tests/test_gpu_parity_clutter.py -m gpu -x -q -s > gpurun_out/pytest_c5_$TAG.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_c5_$TAG.log | tail -8
timeout 900 python bench.py --workload wam_clutter200_1250k --steps 5 --warmup 3 > gpurun_out/bench_c5_$TAG.json 2> gpurun_out/bench_c5_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_c5_$TAG.json; tail -3 gpurun_out/bench_c5_$TAG.err
