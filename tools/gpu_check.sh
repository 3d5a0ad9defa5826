#!/bin/bash
# Mid-session visit: all GPU tests, smoke, the default bench line (with the CPU arm and the parity stats).
set -u
mkdir -p gpurun_out
TAG=${1:-chk}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for w in ${WORKLOADS:-pr2_pr2test1_1m wam_lab1_100k}; do
timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"; tail -3 gpurun_out/bench_${w}_$TAG.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_${w}_$TAG.json').read().strip().splitlines()[-1]); print('  ', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'parity', d.get('parity'))"
done
