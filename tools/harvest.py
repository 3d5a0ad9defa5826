#!/usr/bin/env python
"""Copy a closing visit (tools/gpu_final.sh TAG) from gpurun_out/ into profiles/ and refresh the figures bench.py reads.

  python tools/harvest.py r01D
"""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WORKLOADS = ["pr2_pr2test1_1m", "wam_lab1_100k", "hironx_table_1m", "wam_clutter200_1250k", "wam_wamtest1_birrt100", "wam_lab1_distance_100k"]


def last_line(src, dst):
    lines = [l for l in open(src).read().strip().splitlines() if l.startswith("{")]
    open(dst, "w").write(lines[-1] + "\n")


def main():
    t = sys.argv[1]
    for w in WORKLOADS:
        last_line(os.path.join(G, "bench_%s_%s.json" % (w, t)), os.path.join(P, "%s_bench_%s.json" % (t, w)))
    last_line(os.path.join(G, "bench_ref_%s.json" % t), os.path.join(P, "%s_bench_ref.json" % t))
    shutil.copy(os.path.join(G, "launches_%s.csv" % t), os.path.join(P, "%s_launches.csv" % t))
    shutil.copy(os.path.join(G, "pytest_gpu_%s.log" % t), os.path.join(P, "%s_pytest_gpu.log" % t))
    for k in ("trav", "fk", "trav_clutter"):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                               os.path.join(G, "prof_%s_%s.ncu-rep" % (k, t)), os.path.join(P, "%s_%s" % (t, k))])
    md = open(os.path.join(P, "%s_trav_summary.md" % t)).read()

    def g(k):
        return float(re.search(r"\| %s \| ([0-9.]+)" % re.escape(k), md).group(1))

    phys = {"workload": "pr2_pr2test1_1m",
            "kernel": "cc_traverse_kernel (about 90 %% of the step's GPU time, profiles/%s_launches.csv)" % t,
            "source": "profiles/%s_trav_summary.md (ncu --set full --clock-control none, one 125k-configuration chunk)" % t,
            "issue_active_pct_of_peak": g("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
            "issue_active_pct_while_active": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "fma_pipe_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": g("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
            "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "threads_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "bound_by": "instruction issue + L1/L2 latency (long-scoreboard stalls); the scene is L2-resident, DRAM is idle"}
    json.dump(phys, open(os.path.join(P, "physical_pr2_pr2test1_1m.json"), "w"), indent=1)
    a = json.load(open(os.path.join(P, "%s_trav_traffic.json" % t)))["dram_bytes_per_launch"]
    b = json.load(open(os.path.join(P, "%s_fk_traffic.json" % t)))["dram_bytes_per_launch"]
    traffic = {"workload": "pr2_pr2test1_1m",
               "source": "profiles/%s_trav_traffic.json + profiles/%s_fk_traffic.json (ncu --set full at HEAD, one 125k-config chunk each)" % (t, t),
               "chunks_per_step": 8, "trav_dram_bytes_per_launch": a, "fk_dram_bytes_per_launch": b, "dram_bytes_per_step": 8 * (a + b),
               "note": "physical DRAM bytes of one bench step = 8 chunks x (cc_fk_kernel + cc_traverse_kernel); the link transforms the FK "
                       "kernel writes are what the traversal kernel reads back, the BVH itself stays in L2"}
    json.dump(traffic, open(os.path.join(P, "traffic_pr2_pr2test1_1m.json"), "w"), indent=1)
    for w in WORKLOADS:
        d = json.load(open(os.path.join(P, "%s_bench_%s.json" % (t, w))))
        print("%-26s %12.1f %s  e2e %12.1f  cpu %s  roofline %s  parity %s" % (
            w, d["value"], d["unit"], d["e2e"]["value"], d.get("cpu_baseline", {}).get("value"),
            (d.get("roofline") or {}).get("frac"), (d.get("parity") or {}).get("differing_outside_band")))


if __name__ == "__main__":
    main()
