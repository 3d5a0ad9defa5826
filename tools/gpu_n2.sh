#!/bin/bash
# Two-GPU visit: the bench contract under torchrun (product arm + reference arm), shard-vs-single equality.
set -u
mkdir -p gpurun_out
TAG=${1:-r01e}
N=${2:-2}
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/smi_n${N}_$TAG.csv 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench n=$N rc=$?"
tail -c 900 gpurun_out/bench_n${N}_$TAG.json; tail -5 gpurun_out/bench_n${N}_$TAG.err
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "bench n=1 rc=$?"
tail -c 600 gpurun_out/bench_n1_$TAG.json
