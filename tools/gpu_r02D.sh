#!/bin/bash
# Round-2 visit D: FK captures of C1 / C4 / C5 again, 24 launches each with only the counters the physical / traffic records need
# (the first launches are bind-time probes and certificate grids of a few configurations; tools/harvest_r02.py keeps the largest
# grid; a --set full capture of 24 launches exceeds what gpurun copies back).
set -u
mkdir -p gpurun_out
TAG=${1:-r02C}
M=gpu__time_duration.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
for w in wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:cc_fk -s 2 -c 24 -f -o gpurun_out/${TAG}_${w}_prof_fk \
      python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_${w}_ncu_full_fk.log 2>&1; echo "$w ncu fk rc=$?"; ls -la gpurun_out/${TAG}_${w}_prof_fk.ncu-rep
done
