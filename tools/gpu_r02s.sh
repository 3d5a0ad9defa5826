#!/bin/bash
# Round-2 closing visit: the whole GPU suite at HEAD, smoke, full-size parity sweeps against the unmodified PQP, the bench line of
# every workload (with cpu_baseline + parity) and the reference arm, ncu launch list + full captures of the four kernels of the
# default step, per-chunk timeline, compute-sanitizer over every kernel flavour.  tools/harvest_r02.py TAG turns the captures into
# the files under profiles/ that bench.py reads.
set -u
mkdir -p gpurun_out
TAG=${1:-r02s}
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python tests/parity_sweep.py --workload all --out-dir gpurun_out --tag ${TAG} > gpurun_out/${TAG}_parity_sweep.log 2>&1; echo "parity sweep rc=$?"; tail -8 gpurun_out/${TAG}_parity_sweep.log | cut -c1-300
for w in pr2_pr2test1_1m wam_lab1_100k hironx_table_1m wam_clutter200_1250k wam_wamtest1_birrt100 wam_lab1_distance_100k; do
  [ -f gpurun_out/${TAG}_parity_$w.json ] && cp gpurun_out/${TAG}_parity_$w.json profiles/r02_parity_$w.json     # what bench.py reports as full_size_sweep
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_pr2_pr2test1_1m.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
for w in wam_lab1_100k hironx_table_1m wam_clutter200_1250k wam_wamtest1_birrt100 wam_lab1_distance_100k; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_$w.json 2>> gpurun_out/${TAG}_bench.err; echo "bench $w rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    cb = d.get("cpu_baseline") or {}
    par = cb.get("parity") or {}
    print(f.split("bench_")[1][:-5], round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "cpu", cb.get("value"), "roofline", (d.get("roofline") or {}).get("frac"),
          "physical", {k: (d.get("roofline") or {}).get("physical", {}).get(k) for k in ("issue_frac", "lane_efficiency", "dram_frac")},
          "parity", {k: par.get(k) for k in ("compared", "identical", "near_contact", "mismatch_outside_band")})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_traverse -s 6 -c 1 -f -o gpurun_out/${TAG}_prof_trav \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_trav.log 2>&1; echo "ncu trav rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_fk -s 14 -c 1 -f -o gpurun_out/${TAG}_prof_fk \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_fk.log 2>&1; echo "ncu fk rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cc_broad -s 14 -c 1 -f -o gpurun_out/${TAG}_prof_broad \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full_broad.log 2>&1; echo "ncu broad rc=$?"
timeout 300 python tools/gpu_timeline.py pr2_pr2test1_1m > gpurun_out/${TAG}_timeline_pr2.json 2> gpurun_out/err_tl.txt; tail -2 gpurun_out/err_tl.txt
for v in "" "CC_TUNE_SPLIT=1"; do
  env $v timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/gpu_sanitize.py > gpurun_out/${TAG}_sanitize_memcheck_${v:-default}.log 2>&1; echo "memcheck ${v:-default} rc=$?"; grep -E "ERROR SUMMARY|ok" gpurun_out/${TAG}_sanitize_memcheck_${v:-default}.log | tail -8
done
CC_TUNE_SPLIT=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/gpu_sanitize.py > gpurun_out/${TAG}_sanitize_racecheck_CC_TUNE_SPLIT=1.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ok" gpurun_out/${TAG}_sanitize_racecheck_CC_TUNE_SPLIT=1.log | tail -8
