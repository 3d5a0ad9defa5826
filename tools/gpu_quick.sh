#!/bin/bash
# Quick visit after a kernel change: parity tests under short timeouts, then the benches without the CPU baseline.
set -u
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_quick_$TAG.log 2>&1; rc=$?; echo "wam pytest rc=$rc"; tail -3 gpurun_out/pytest_quick_$TAG.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "full pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for w in ${WORKLOADS:-pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k}; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2))
"; tail -2 gpurun_out/err_$TAG.txt
done
if [ "${PROBE:-0}" = "1" ]; then bash tools/gpu_probe3.sh $TAG; fi
