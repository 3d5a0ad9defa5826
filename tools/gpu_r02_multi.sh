#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the in-library multi-device scene (shards == single, bitwise), then the bench contract under
# torchrun at N GPUs (device-resident value with one ncclAllGather; e2e through ONE process driving all N devices).
set -u
mkdir -p gpurun_out
N=${1:-2}
TAG=${2:-r02m}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_multi_device.py -m gpu -q > gpurun_out/${TAG}_pytest_multi_n$N.log 2>&1; echo "multi-device pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_multi_n$N.log
[ -n "${SKIP_N1:-}" ] || timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench N=1 rc=$?"; cut -c1-160 gpurun_out/${TAG}_bench_n1.json
for w in ${WORKLOADS:-pr2_pr2test1_1m}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --workload $w --no-cpu-baseline \
      > gpurun_out/${TAG}_bench_n${N}_$w.json 2> gpurun_out/${TAG}_bench_n${N}_$w.err; echo "bench N=$N $w rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench_n${N}_$w.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$w N=$N value %.1f M/s  e2e %.1f M/s  ms %.3f  multi %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["config"].get("multi_device_e2e")))
PY
  tail -3 gpurun_out/${TAG}_bench_n${N}_$w.err
done
