#!/bin/bash
# Round-2 visits x, y: FK2 variants (x: joint transforms in their own parallel phase — slower, reverted; y: structure-of-arrays step constants) forced on / off / by plan.
set -u
mkdir -p gpurun_out
TAG=${1:-r02x}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fk_ladder.py tests/test_affine_dofs.py tests/test_grabbed.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick.log 2>&1; rc=$?; echo "quick pytest rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_quick.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/${TAG}_pytest_quick.log; exit 1; fi
CC_TUNE_FK2=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fk_ladder.py tests/test_affine_dofs.py tests/test_grabbed.py -m gpu -x -q > gpurun_out/${TAG}_pytest_quick_fk2.log 2>&1; rc=$?; echo "quick pytest (FK2 forced) rc=$rc"; tail -3 gpurun_out/${TAG}_pytest_quick_fk2.log
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
for knobs in "X=byplan" "CC_TUNE_FK2=1" "CC_TUNE_FK2=0" "X=byplan"; do echo "=== $knobs"; bench4 $knobs; done 2>&1 | tee gpurun_out/${TAG}_ab_fk2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cc_fk2 -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_fk2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_fk2.log 2>&1; echo "ncu fk2 rc=$?"
timeout 120 python tools/gpu_timeline.py pr2_pr2test1_1m > gpurun_out/${TAG}_timeline_pr2.json 2>> gpurun_out/err_$TAG.txt; python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_timeline_pr2.json"))
s = d["steps"][1]
print("timeline step", s["step_ms"], [(c["lane"], c["lane_free_ms"], c["fk_end_ms"], c["traversal_end_ms"]) for c in s["chunks"]])
PY
