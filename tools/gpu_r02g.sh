#!/bin/bash
# Round-2 visit g: A/B of the broadphase reading link boxes from shared memory (CC_BROAD_SMEM), parity on the new default,
# A/B of the distance kernel's push order, the FK kernel's ncu capture (a real 2048-block launch this time).
set -u
mkdir -p gpurun_out
TAG=${1:-r02g}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
for lib in nobroadsmem base nobroadsmem base; do
  echo "=== $lib"
  for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
    CUDACOLLISION_LIB=$PWD/build/libcc_$lib.so timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$w', round(d['value']/1e6, 2), 'M/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value']/1e6, 2), 'first-hit', round(d['config']['first_hit_mode_configs_per_s']/1e6, 2), 'dynamic', round(d['config']['dynamic_scene_configs_per_s']/1e6, 2))
"; tail -2 gpurun_out/err_$TAG.txt
  done
done 2>&1 | tee gpurun_out/${TAG}_ab_broad_smem.log
for lib in base distrev base distrev; do
  echo "=== $lib"
  CUDACOLLISION_LIB=$PWD/build/libcc_$lib.so timeout 300 python bench.py --workload wam_lab1_distance_100k --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/err_$TAG.txt | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('distance exact', round(d['value']/1e6, 3), 'M/s', round(d['ms_per_step'], 3), 'ms; with pqprave tolerances', round(d['config']['configs_per_s_with_pqprave_tolerances']/1e6, 3), 'M/s')
"; tail -2 gpurun_out/err_$TAG.txt
done 2>&1 | tee gpurun_out/${TAG}_ab_dist_push.log
CUDACOLLISION_LIB=$PWD/build/libcc_distrev.so timeout 300 python -m pytest tests/test_distance.py tests/test_pair_tables.py -m gpu -x -q 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_fk -s 14 -c 1 -f -o gpurun_out/${TAG}_prof_fk \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_fk.log 2>&1; echo "ncu fk rc=$?"
