#!/bin/bash
# A/B visit: for each library variant run the parity tests (quick subset) and the probe.
set -u
mkdir -p gpurun_out
TAG=${1:-ab}
shift
for lib in "$@"; do
  echo "=== $lib"
  export CUDACOLLISION_LIB=$PWD/$lib
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}_$(basename $lib .so).log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_${TAG}_$(basename $lib .so).log
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-220
  timeout 300 python tools/gpu_probe.py 100000 2>&1 | grep -v "counters=0" | head -8
done
