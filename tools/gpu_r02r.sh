#!/bin/bash
# Round-2 visit r: FK kernel form chosen per plan; chunking of mid-size batches (one chunk = no overlap at all).
set -u
mkdir -p gpurun_out
TAG=${1:-r02r}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fk_ladder.py tests/test_reference_collision_tests.py tests/test_edge_check.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick.log 2>&1; rc=$?; echo "quick pytest rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_quick.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/${TAG}_pytest_quick.log; exit 1; fi
bench4() {
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
    env "$@" timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2), 'dynamic', round(d['config']['dynamic_scene_configs_per_s']/1e6, 2), 'launches', d['gpu_launches'])
"; tail -2 gpurun_out/err_$TAG.txt
  done
}
for knobs in "X=base"; do echo "=== $knobs"; bench4 $knobs; done 2>&1 | tee gpurun_out/${TAG}_ab.log
for c in 0 25088 33408 50048 66688 100096 131072; do
  for w in wam_lab1_100k pr2_pr2test1_1m; do
    echo "--- CC_TUNE_CHUNK=$c $w"; CC_TUNE_CHUNK=$c timeout 200 python tools/gpu_scale_n.py $w 50000,100000,200000,400000 2>&1 | grep "n="
  done
done 2>&1 | tee gpurun_out/${TAG}_chunk_sweep.log
