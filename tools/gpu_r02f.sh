#!/bin/bash
# Round-2 visit f: new GPU tests (affine DOFs, rebound constraint function, multi-device), then ncu: launch list + a full capture
# of the traversal kernel with sources (one 256k-configuration chunk of the PR2 bench step).
set -u
mkdir -p gpurun_out
TAG=${1:-r02f}
timeout 900 python -m pytest tests/test_affine_dofs.py tests/test_edge_check.py tests/test_multi_device.py tests/test_refresh.py tests/test_fk_ladder.py -m gpu -x -q > gpurun_out/${TAG}_pytest_new.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_new.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_traverse -s 6 -c 1 -f -o gpurun_out/${TAG}_prof_trav \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_trav.log 2>&1; echo "ncu trav rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_fk -s 6 -c 1 -f -o gpurun_out/${TAG}_prof_fk \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_fk.log 2>&1; echo "ncu fk rc=$?"
ls -la gpurun_out | grep $TAG
