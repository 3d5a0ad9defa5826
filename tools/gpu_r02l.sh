#!/bin/bash
# Round-2 visit l: cc_broad_kernel restructured (8 configurations per warp, the pair list walked per configuration so the boxes hit
# L1) against the in-warp broadphase; select-built descent outputs; FK kernel at 5 / 4 resident blocks (no spilled loop counter).
set -u
mkdir -p gpurun_out
TAG=${1:-r02l}
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_pair_tables.py tests/test_analytic_primitives.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick.log 2>&1; rc=$?; echo "quick pytest rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_quick.log
if [ $rc -ne 0 ]; then tail -40 gpurun_out/${TAG}_pytest_quick.log; exit 1; fi
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
for knobs in "CC_TUNE_CAND=0" "CC_TUNE_CAND=1" "CUDACOLLISION_LIB=$PWD/build/libcc_outsel.so" "CUDACOLLISION_LIB=$PWD/build/libcc_fk5.so" "CUDACOLLISION_LIB=$PWD/build/libcc_fk4.so" "CC_TUNE_CAND=1"; do echo "=== $knobs"; bench4 $knobs; done 2>&1 | tee gpurun_out/${TAG}_ab.log
timeout 120 python tools/gpu_timeline.py pr2_pr2test1_1m > gpurun_out/${TAG}_timeline_pr2.json 2>> gpurun_out/err_$TAG.txt; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02l_timeline_pr2.json"))
s = d["steps"][1]
print("timeline step", s["step_ms"], [(c["lane"], c["lane_free_ms"], c["fk_end_ms"], c["traversal_end_ms"]) for c in s["chunks"]])
PY
