#!/bin/bash
# Round-2 visit C: ncu launch lists + full captures of the three kernels for the other collision workloads (C1, C4, C5), so that
# their bench lines carry roofline.physical / roofline.traffic too (tools/harvest_r02.py r02C_<workload> <chunk> <workload> <n>);
# compute-sanitizer initcheck over the round-2 flavours.
set -u
mkdir -p gpurun_out
TAG=${1:-r02C}
for w in wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_${w}_launches.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${w}_ncu_launch.log 2>&1; echo "$w ncu launches rc=$?"
  for k in trav:cc_traverse:6 fk:cc_fk:6 broad:cc_broad:6; do
    IFS=: read name regex skip <<< "$k"
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o gpurun_out/${TAG}_${w}_prof_$name \
        python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${w}_ncu_full_$name.log 2>&1; echo "$w ncu $name rc=$?"
  done
done
CC_TUNE_SPLIT=1 timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/gpu_sanitize.py > gpurun_out/${TAG}_sanitize_initcheck_CC_TUNE_SPLIT=1.log 2>&1; echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|ok" gpurun_out/${TAG}_sanitize_initcheck_CC_TUNE_SPLIT=1.log | tail -8
