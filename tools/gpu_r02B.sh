#!/bin/bash
# Round-2 visit B: 2 warps per traversal block (14 blocks per SM: slots free sooner in a chunk's tail) and the triangle-queue
# thresholds re-checked now that the broadphase has left the traversal kernel.
set -u
mkdir -p gpurun_out
TAG=${1:-r02B}
bench4() {
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
    env "$@" timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2), 'dynamic', round(d['config']['dynamic_scene_configs_per_s']/1e6, 2))
"; tail -2 gpurun_out/err_$TAG.txt
  done
}
for knobs in "X=base" "CUDACOLLISION_LIB=$PWD/build/libcc_w2.so" "CUDACOLLISION_LIB=$PWD/build/libcc_tri32.so" "CUDACOLLISION_LIB=$PWD/build/libcc_tri8.so" "X=base"; do echo "=== $knobs"; bench4 $knobs; done 2>&1 | tee gpurun_out/${TAG}_ab.log
