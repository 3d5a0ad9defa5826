#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-b3}
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/pytest_gpu_$TAG.log | tail -12
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err; echo "bench $w rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${w}_$TAG.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline",{}).get("value"), "roof", d.get("roofline",{}).get("frac"), {k:d["config"][k] for k in ("pairs_env","pairs_self","pairs_hoisted","colliding_fraction")})
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_${w}_$TAG.err").read()[-1500:])
PY
done
