#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r01n}
timeout 900 python -m pytest tests/test_pair_tables.py -m gpu -x -q -s > gpurun_out/pytest_tables_$TAG.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_tables_$TAG.log
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
for t in on off; do
  if [ $t = off ]; then export CC_TUNE_NO_TABLES=1; else unset CC_TUNE_NO_TABLES; fi
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w tables=$t', round(d['value']/1e6, 2), 'M/s  e2e', round(d['e2e']['value']/1e6, 2))
"
done; done
