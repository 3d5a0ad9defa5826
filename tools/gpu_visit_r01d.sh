#!/bin/bash
# Round-1 closing visit: parity tests at HEAD, the four bench workloads (+ reference arm), ncu launch list and full captures.
set -u
mkdir -p gpurun_out
TAG=${1:-r01d}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.csv 2>&1
nproc > gpurun_out/nproc_$TAG.txt; grep -m1 'model name' /proc/cpuinfo >> gpurun_out/nproc_$TAG.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_gpu_$TAG.log | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"
  tail -c 600 gpurun_out/bench_${w}_$TAG.json; echo
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
tail -c 700 gpurun_out/bench_ref_$TAG.json
bash tools/gpu_prof.sh $TAG
ls -la gpurun_out
