#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01g}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_pr2_$TAG.json 2> gpurun_out/bench_pr2_$TAG.err; echo "bench rc=$?"; tail -c 700 gpurun_out/bench_pr2_$TAG.json
