#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01h}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_$TAG.log
python tools/gpu_latency.py 2>&1 | tail -9; cp gpurun_out/latency.json gpurun_out/latency_$TAG.json
CC_TUNE_SPLIT=1 python tools/gpu_latency.py 2>&1 | tail -9; cp gpurun_out/latency.json gpurun_out/latency_nosplit_$TAG.json
timeout 600 python bench.py --workload wam_wamtest1_birrt100 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_birrt_$TAG.json 2> gpurun_out/bench_birrt_$TAG.err; echo "bench birrt rc=$?"
grep -o '"value": [0-9.]*' gpurun_out/bench_birrt_$TAG.json | head -1
CC_TUNE_SPLIT=1 timeout 600 python bench.py --workload wam_wamtest1_birrt100 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -o '"value": [0-9.]*' | head -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -o '"value": [0-9.]*' | head -1
