#!/bin/bash
# last visit of the round: the external-mimic test, the parity / certificate / FK suites on the changed hoisting rule, the default bench
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_external_mimic.py tests/test_gpu_parity.py tests/test_pair_tables.py tests/test_fk_ladder.py -m gpu -x -q > gpurun_out/r02K_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02K_tests.log
tail -5 gpurun_out/r02K_tests.log
timeout 80 python bench.py --steps 10 --warmup 3 > gpurun_out/r02K_bench.json 2> gpurun_out/r02K_bench.err
echo "bench rc=$?"
cut -c1-400 gpurun_out/r02K_bench.json
