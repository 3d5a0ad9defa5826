#!/bin/bash
# Development visit: parity tests first (under a timeout), then a short bench and the device work counters.
set -u
mkdir -p gpurun_out
TAG=${1:-dev}
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
if [ "${SANITIZE:-0}" = "1" ]; then
  timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "empty or pose_mode or hoisted" > gpurun_out/sanitizer_$TAG.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer_$TAG.log
fi
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python tools/gpu_probe.py 100000 > gpurun_out/probe_$TAG.log 2>&1; cat gpurun_out/probe_$TAG.log
