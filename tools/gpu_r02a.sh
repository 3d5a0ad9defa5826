#!/bin/bash
# Round-2 first visit: GPU tests, the full-size parity sweeps, the default bench line, then timing experiments
# (FK share, chunk size, per-chunk timeline, work trace).  Everything lands in gpurun_out/ under the tag.
set -u
mkdir -p gpurun_out
TAG=${1:-r02a}
nproc > gpurun_out/${TAG}_host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/${TAG}_host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 1500 python tests/parity_sweep.py --workload all --out-dir gpurun_out --tag $TAG > gpurun_out/${TAG}_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/${TAG}_parity.log | cut -c1-600
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench.json
for knobs in "X=0" "CC_TUNE_SKIP_FK=1" "CC_TUNE_CHUNK=62464" "CC_TUNE_CHUNK=250112" "CC_TUNE_CHUNK=500096" "CC_TUNE_CHUNK=1000064" "CC_TUNE_NO_TABLES=1"; do
  echo "=== $knobs"
  for w in pr2_pr2test1_1m wam_lab1_100k; do
    env $knobs timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2))
"; tail -2 gpurun_out/err_$TAG.txt
  done
done 2>&1 | tee gpurun_out/${TAG}_knobs.log
timeout 300 python tools/gpu_timeline.py pr2_pr2test1_1m > gpurun_out/${TAG}_timeline_pr2.json 2> gpurun_out/err_tl.txt; tail -2 gpurun_out/err_tl.txt
timeout 300 python tools/gpu_probe.py 1000000 pr2_pr2test1_1m > gpurun_out/${TAG}_probe_pr2.log 2>&1; tail -12 gpurun_out/${TAG}_probe_pr2.log
