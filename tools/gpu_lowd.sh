#!/bin/bash
set -u
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m; do
for n in none 1 2 3; do
  if [ $n = none ]; then unset CC_EXPERIMENT_DROP_LOWD; else export CC_EXPERIMENT_DROP_LOWD=$n; fi
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w drop<=$n', round(d['value']/1e6, 2), 'M/s', 'pairs', d['config']['pairs_env'], d['config']['pairs_self'], d['config']['pairs_hoisted'], 'colliding', d['config']['colliding_fraction'])
"
done; done
