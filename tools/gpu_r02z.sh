#!/bin/bash
# Round-2 visit z: host-buffer batches start with a quarter and a half chunk (CC_TUNE_RAMP=0: equal chunks).
set -u
mkdir -p gpurun_out
TAG=${1:-r02z}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick.log 2>&1; rc=$?; echo "quick pytest rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_quick.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/${TAG}_pytest_quick.log; exit 1; fi
for knobs in "CC_TUNE_RAMP=1" "CC_TUNE_RAMP=0" "CC_TUNE_RAMP=1" "CC_TUNE_RAMP=0"; do echo "=== $knobs";
  for w in pr2_pr2test1_1m hironx_table_1m wam_clutter200_1250k; do
    env $knobs timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2))
"; tail -2 gpurun_out/err_$TAG.txt
  done
done 2>&1 | tee gpurun_out/${TAG}_ab_ramp.log
