#!/bin/bash
# BiRRT visit: GPU parity of the planner, the plans/s bench workload and its reference arm.
set -u
mkdir -p gpurun_out
TAG=${1:-r01f}
timeout 900 python -m pytest tests/test_birrt.py tests/test_edge_check.py -m gpu -x -q -s > gpurun_out/pytest_birrt_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_birrt_$TAG.log
timeout 600 python bench.py --workload wam_wamtest1_birrt100 --steps 5 --warmup 3 > gpurun_out/bench_birrt_$TAG.json 2> gpurun_out/bench_birrt_$TAG.err; echo "bench birrt rc=$?"
tail -c 1500 gpurun_out/bench_birrt_$TAG.json; tail -3 gpurun_out/bench_birrt_$TAG.err
timeout 600 python bench.py --impl reference --workload wam_wamtest1_birrt100 --steps 2 --warmup 1 > gpurun_out/bench_birrt_ref_$TAG.json 2> gpurun_out/bench_birrt_ref_$TAG.err; echo "ref rc=$?"
tail -c 500 gpurun_out/bench_birrt_ref_$TAG.json
nproc
