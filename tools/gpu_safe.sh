#!/bin/bash
# First contact with risky kernel changes: short timeouts so that a hang cannot become a strike.
set -u
mkdir -p gpurun_out
TAG=${1:-safe}
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "empty or determinism or verdicts" > gpurun_out/pytest_safe_$TAG.log 2>&1; rc=$?; echo "quick pytest rc=$rc"; tail -5 gpurun_out/pytest_safe_$TAG.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "full pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2))
"; tail -2 gpurun_out/err_$TAG.txt
done
timeout 200 python tools/gpu_probe.py 100000 2>&1 | tail -5
