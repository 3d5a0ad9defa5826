#!/usr/bin/env python
"""Summarise an ncu report (run here, no GPU needed): key counters, stall mix, hottest SASS lines.

  python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/X   ->  profiles/X_summary.md, profiles/X_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__t_sectors.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    lines = ["# ncu summary of `%s`" % rep.split("/")[-1], "",
             "Captured with `ncu --set full --clock-control none --import-source on` (cold-cache, serialised replays; use for",
             "ratios and shares, not for absolute time).", ""]
    traffic = {}
    for li, r in enumerate(raw[2:]):
        kname = r[hdr.index("Kernel Name")]
        lines += ["## launch %d: `%s`" % (li, kname), "", "| metric | value | unit |", "|---|---|---|"]
        vals = dict(zip(hdr, r))
        un = dict(zip(hdr, units))
        for k in KEYS:
            if k in vals:
                lines.append("| %s | %s | %s |" % (k, vals[k], un[k]))
        stall = sorted(((float(vals[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and
                        h.endswith("_per_issue_active.ratio") and vals[h]), reverse=True)
        lines += ["", "Warp-stall mix (warps stalled per issue-active cycle):", ""]
        for v, h in stall[:8]:
            lines.append("* %s: %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
        lines.append("")

        def tobytes(key):
            v, u = float(vals[key]), un[key].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        rd, wr = tobytes("dram__bytes_read.sum"), tobytes("dram__bytes_write.sum")
        grid = int(float(vals.get("launch__grid_size", "0") or 0))
        want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
        if traffic and (traffic.get("grid_size") == want or (grid != want and grid <= traffic.get("grid_size", -1))):
            continue                      # several launches captured: the one with the wanted grid (argv[3]), else the largest
        traffic = {"grid_size": grid, "launch_index": li, "kernel": kname, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                   "gpu_time_ms_under_ncu": float(vals["gpu__time_duration.sum"]) * (1e-3 if un["gpu__time_duration.sum"] == "us" else 1.0)
                   if un["gpu__time_duration.sum"] in ("ms", "us") else vals["gpu__time_duration.sum"],
                   "l2_hit_rate_pct": float(vals.get("lts__t_sector_hit_rate.pct", "nan")), "source": rep.split("/")[-1]}
    src = ncu_csv(rep, "source")
    if len(src) > 2 and "# Samples" in src[1]:
        h = src[1]
        d = src[2:]
        isamp, iex, ith, isrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Avg. Threads Executed"), h.index("Source")
        tot = sum(int(r[isamp]) for r in d if len(r) > isamp and r[isamp].isdigit()) or 1
        lines += ["## hottest SASS lines of the last captured launch (by stall samples)", "",
                  "| line | samples % | executed | avg threads | SASS |", "|---|---|---|---|---|"]
        idx = sorted(range(len(d)), key=lambda i: -int(d[i][isamp]) if d[i][isamp].isdigit() else 0)[:30]
        for i in sorted(idx):
            r = d[i]
            lines.append("| %d | %.2f | %s | %s | `%s` |" % (i, 100.0 * int(r[isamp]) / tot, r[iex], r[ith], r[isrc].strip()[:80]))
        lines.append("")
    open(prefix + "_summary.md", "w").write("\n".join(lines))
    json.dump(traffic, open(prefix + "_traffic.json", "w"), indent=1)
    print("wrote", prefix + "_summary.md", prefix + "_traffic.json")


if __name__ == "__main__":
    main()
