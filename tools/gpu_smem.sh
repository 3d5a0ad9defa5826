#!/bin/bash
set -u
mkdir -p gpurun_out
export CUDACOLLISION_LIB=$PWD/build/libcc_x.so
for cfg in "7 100" "6 80" "5 68" "4 55" "3 42"; do
  set -- $cfg
  export CC_TUNE_BLOCKS_PER_SM=$1 CC_TUNE_CARVEOUT=$2
  echo "=== blocks/SM $1 carveout $2%"
  for w in pr2_pr2test1_1m hironx_table_1m; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms')
"
  done
done
