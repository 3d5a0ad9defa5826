#!/bin/bash
# Round-2 visit c: folded FK chain program + slot-major pose scratch.  Parity first, then the 2 x 2 A/B, then the timeline.
set -u
mkdir -p gpurun_out
TAG=${1:-r02c}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
for knobs in "CC_TUNE_FK_FOLD=0 CC_TUNE_SLOT_MAJOR=0" "CC_TUNE_FK_FOLD=1 CC_TUNE_SLOT_MAJOR=0" "CC_TUNE_FK_FOLD=0 CC_TUNE_SLOT_MAJOR=1" "CC_TUNE_FK_FOLD=1 CC_TUNE_SLOT_MAJOR=1" "CC_TUNE_SKIP_FK=1"; do
  echo "=== $knobs"
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
    env $knobs timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2), 'launches', d['gpu_launches'])
"; tail -2 gpurun_out/err_$TAG.txt
  done
done 2>&1 | tee gpurun_out/${TAG}_ab_fk.log
timeout 300 python tools/gpu_timeline.py pr2_pr2test1_1m > gpurun_out/${TAG}_timeline_pr2.json 2> gpurun_out/err_tl.txt; tail -2 gpurun_out/err_tl.txt
python - <<'PY'
import json
t = json.load(open("gpurun_out/%s_timeline_pr2.json" % "r02c"))
st = t["steps"][0]
print("step", round(st["step_ms"], 3), "fk sum", round(st["sum_of_fk_intervals_ms"], 3), "chunks", [(c["lane"], c["lane_free_ms"], c["fk_end_ms"], c["traversal_end_ms"]) for c in st["chunks"]])
PY
