#!/bin/bash
# A/B visit: for each library variant (paths relative to the repo root) run the benches without the CPU baseline.
# PYTEST=1 also runs the parity tests for each variant.
set -u
mkdir -p gpurun_out
TAG=${1:-ab}
shift
for lib in "$@"; do
  echo "=== $lib"
  export CUDACOLLISION_LIB=$PWD/$lib
  if [ "${PYTEST:-0}" = "1" ]; then
    timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}_$(basename $lib .so).log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_${TAG}_$(basename $lib .so).log
  fi
  for w in ${WORKLOADS:-pr2_pr2test1_1m wam_lab1_100k hironx_table_1m}; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2))
"; tail -2 gpurun_out/err_$TAG.txt
  done
done 2>&1 | tee gpurun_out/ab_$TAG.log
