#!/bin/bash
# Round-2 visit j: GPU suite at HEAD, smoke, the default bench line with cpu_baseline + parity, the reference arm, the other
# workloads' full lines, refresh latency.
set -u
mkdir -p gpurun_out
TAG=${1:-r02j}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_pr2_pr2test1_1m.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02j_bench_pr2_pr2test1_1m.json") if l.startswith("{")][-1])
print("value %.2f M/s e2e %.2f M/s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6))
print("roofline frac %.2f physical %s" % (d["roofline"]["frac"], {k: d["roofline"]["physical"].get(k) for k in ("issue_frac", "lane_efficiency", "dram_frac")}))
print("parity", {k: v for k, v in d["cpu_baseline"]["parity"].items() if k != "full_size_sweep"})
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
for w in wam_lab1_100k hironx_table_1m wam_clutter200_1250k wam_wamtest1_birrt100 wam_lab1_distance_100k; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2>> gpurun_out/${TAG}_bench.err; echo "bench $w rc=$?"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/${TAG}_bench_$w.json') if l.startswith('{')][-1]); print('  ', round(d['value'],1), d['unit'], 'e2e', round(d['e2e']['value'],1), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'parity', {k: v for k, v in ((d.get('cpu_baseline') or {}).get('parity') or {}).items() if k in ('compared','identical','mismatch_outside_band','near_contact')})"
done
timeout 600 python tools/gpu_refresh_latency.py pr2_pr2test1_1m > gpurun_out/${TAG}_refresh_latency_pr2.json 2> gpurun_out/err_$TAG.txt; tail -2 gpurun_out/err_$TAG.txt; cat gpurun_out/${TAG}_refresh_latency_pr2.json
