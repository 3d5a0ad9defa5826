#!/bin/bash
# Round-2 visit q: two-phase FK kernel (cc_fk2_kernel) against the one-thread-per-configuration form; packed 15-bit box tests in
# cc_broad_kernel; 6 resident traversal blocks per SM (room for the FK / broadphase blocks of the next chunk); lanes.
set -u
mkdir -p gpurun_out
TAG=${1:-r02q}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fk_ladder.py tests/test_pair_tables.py tests/test_refresh.py tests/test_affine_dofs.py tests/test_grabbed.py tests/test_distance.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick.log 2>&1; rc=$?; echo "quick pytest rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_quick.log
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
for knobs in "X=base" "CC_TUNE_FK2=0" "CC_TUNE_BLOCKS_PER_SM=6" "CC_TUNE_LANES=2" "CC_TUNE_LANES=3" "X=base"; do echo "=== $knobs"; bench4 $knobs; done 2>&1 | tee gpurun_out/${TAG}_ab.log
timeout 120 python tools/gpu_timeline.py pr2_pr2test1_1m > gpurun_out/${TAG}_timeline_pr2.json 2>> gpurun_out/err_$TAG.txt; python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_timeline_pr2.json"))
s = d["steps"][1]
print("timeline step", s["step_ms"], [(c["lane"], c["lane_free_ms"], c["fk_end_ms"], c["traversal_end_ms"]) for c in s["chunks"]])
PY
timeout 300 python bench.py --workload wam_wamtest1_birrt100 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-160
timeout 300 python bench.py --workload wam_lab1_distance_100k --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
