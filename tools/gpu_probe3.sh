#!/bin/bash
# Work counters of the three robot workloads (no timing claims: counters on).
set -u
mkdir -p gpurun_out
TAG=${1:-probe}
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m; do
  timeout 300 python tools/gpu_probe.py 100000 $w 2>&1 | grep -v "counters=0" | tee -a gpurun_out/probe_$TAG.log
done
