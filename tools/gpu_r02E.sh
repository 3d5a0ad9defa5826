mkdir -p gpurun_out
for w in wam_lab1_100k hironx_table_1m wam_clutter200_1250k; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r02s_bench_$w.json 2>> gpurun_out/r02s_bench3.err; echo "bench $w rc=$?"
done
