#!/usr/bin/env python
"""Turn the ncu captures of a round-2 profiling visit (tools/gpu_r02f.sh TAG) into the committed evidence bench.py reads:

  profiles/<TAG>_trav_summary.md / _traffic.json, <TAG>_fk_summary.md / _traffic.json   (tools/ncu_summary.py)
  profiles/<TAG>_launches.csv                                                              (the launch list of the bench command)
  profiles/<TAG>_trav_sass.txt                                                             (SASS of the captured traversal kernel)
  profiles/physical_<workload>.json, profiles/traffic_<workload>.json                     (what bench.py's roofline.physical uses)

  python tools/harvest_r02.py r02f [configs_per_launch]
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def raw_metrics(rep, want_grid=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    gi = rows[0].index("launch__grid_size")
    # several launches captured (bind-time probes, certificate grids, the bench chunks): the one whose grid is the bench chunk's,
    # else the largest
    match = [r for r in rows[2:] if want_grid and int(float(r[gi] or 0)) == want_grid]
    best = match[0] if match else max(rows[2:], key=lambda r: float(r[gi] or 0))
    return dict(zip(rows[0], best))


def main():
    # harvest_r02.py TAG [configs_per_launch] [workload] [configs_per_step]; the captures of a workload other than the default are
    # named TAG_<workload>_prof_*.ncu-rep and TAG_<workload>_launches.csv (the TAG passed here is then "TAG_<workload>")
    t = sys.argv[1]
    workload = sys.argv[3] if len(sys.argv) > 3 else "pr2_pr2test1_1m"
    shutil.copy(os.path.join(G, "%s_launches.csv" % t), os.path.join(P, "%s_launches.csv" % t))
    has_broad = os.path.isfile(os.path.join(G, "%s_prof_broad.ncu-rep" % t))
    cfg0 = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    for k in ("trav", "fk") + (("broad",) if has_broad else ()):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                               os.path.join(G, "%s_prof_%s.ncu-rep" % (t, k)), os.path.join(P, "%s_%s" % (t, k))] +
                              ([str((cfg0 + 127) // 128)] if k == "fk" else []))
    sass = subprocess.run(["ncu", "-i", os.path.join(G, "%s_prof_trav.ncu-rep" % t), "--page", "source", "--csv", "--print-source", "sass"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    hdr = [r for r in rows if r and r[0] in ("Address", "#")]
    with open(os.path.join(P, "%s_trav_sass.txt" % t), "w") as f:
        f.write("# SASS of the captured cc_traverse_kernel launch with per-instruction execution counts (ncu --page source --print-source sass)\n")
        for r in rows:
            if len(r) > 2:
                f.write("\t".join(r[:6]) + "\n")
    # launch list: configurations per traversal launch = the chunk size of the bench step
    nlaunch = {"trav": 0, "fk": 0, "broad": 0}
    dur = {"trav": 0.0, "fk": 0.0, "broad": 0.0}
    for r in csv.reader(open(os.path.join(P, "%s_launches.csv" % t), errors="ignore")):
        if len(r) > 5 and "cc_" in " ".join(r):
            name = " ".join(r)
            k = "trav" if "cc_traverse" in name else ("fk" if "cc_fk" in name else ("broad" if "cc_broad" in name else None))
            if k:
                nlaunch[k] += 1
                try:
                    dur[k] += float(r[-1])
                except ValueError:
                    pass
    cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    mt = raw_metrics(os.path.join(G, "%s_prof_trav.ncu-rep" % t))
    mf = raw_metrics(os.path.join(G, "%s_prof_fk.ncu-rep" % t), want_grid=(cfg + 127) // 128)      # cc_fk_kernel: 128 configurations per block
    mb = raw_metrics(os.path.join(G, "%s_prof_broad.ncu-rep" % t)) if has_broad else {}
    n_step = int(sys.argv[4]) if len(sys.argv) > 4 else 1000000
    chunks = (n_step + cfg - 1) // cfg
    wt, wf = float(mt["smsp__inst_executed.sum"]), float(mf["smsp__inst_executed.sum"])
    # the captured FK launch may be one of the certificate builder's (another batch size): cc_fk_kernel runs one configuration per
    # thread, 128 per block, so its own size is known; scale its counters to the bench chunk
    fk_scale = 1.0
    if "cc_fk_kernel" in mf.get("Kernel Name", "") and float(mf["launch__grid_size"]) * 128 != cfg:
        fk_scale = cfg / (float(mf["launch__grid_size"]) * 128.0)
        wf *= fk_scale
    wb = float(mb["smsp__inst_executed.sum"]) if has_broad else 0.0
    last = min(cfg, n_step - (chunks - 1) * cfg)

    def g(m, k):
        return float(m[k]) if m.get(k) not in (None, "") else None

    phys = {"workload": workload,
            "kernels": "cc_traverse_kernel + cc_fk_kernel / cc_fk2_kernel + cc_broad_kernel: every kernel of the bench step (profiles/%s_launches.csv: %d + %d + %d launches)" % (
                t, nlaunch["trav"], nlaunch["fk"], nlaunch["broad"]),
            "source": "profiles/%s_trav_summary.md + profiles/%s_fk_summary.md (ncu --set full --clock-control none --import-source on, at HEAD, "
                      "one %d-configuration chunk each); SASS with execution counts in profiles/%s_trav_sass.txt" % (t, t, cfg, t),
            "configurations_per_captured_launch": cfg,
            "fk_capture_note": (None if fk_scale == 1.0 else "the captured cc_fk_kernel launch is one of the certificate builder's (%d configurations, every link "
                                "emitted); its counters are scaled to the bench chunk" % int(round(cfg / fk_scale))),
            "warp_instructions_per_configuration": (wt + wf + wb) / cfg,
            "warp_instructions_per_configuration_broadphase": wb / cfg,
            "warp_instructions_per_configuration_traversal": wt / cfg,
            "warp_instructions_per_configuration_fk": wf / cfg,
            "threads_per_instruction": g(mt, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "issue_active_pct_while_active_traversal": g(mt, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct_of_elapsed_traversal_captured_alone": g(mt, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
            "fma_pipe_pct": g(mt, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": g(mt, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "lsu_pipe_pct": g(mt, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": g(mt, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": g(mt, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": g(mt, "lts__t_sector_hit_rate.pct"),
            "registers_per_thread": g(mt, "launch__registers_per_thread"),
            "bound_by": "instruction issue + L1/L2 latency (long-scoreboard stalls): the scene is L2-resident, DRAM carries only q, the results and the "
                        "per-chunk link poses the FK kernel hands to the traversal kernel"}
    json.dump(phys, open(os.path.join(P, "physical_%s.json" % workload), "w"), indent=1)
    a = json.load(open(os.path.join(P, "%s_trav_traffic.json" % t)))["dram_bytes_per_launch"]
    b = json.load(open(os.path.join(P, "%s_fk_traffic.json" % t)))["dram_bytes_per_launch"] * fk_scale
    cbytes = json.load(open(os.path.join(P, "%s_broad_traffic.json" % t)))["dram_bytes_per_launch"] if has_broad else 0.0
    scale = (chunks - 1) + last / float(cfg)
    traffic = {"workload": workload,
               "source": "profiles/%s_trav_traffic.json + profiles/%s_fk_traffic.json (ncu --set full at HEAD, one %d-configuration chunk each)" % (t, t, cfg),
               "chunks_per_step": chunks, "configurations_per_captured_launch": cfg, "trav_dram_bytes_per_launch": a, "fk_dram_bytes_per_launch": b,
               "broad_dram_bytes_per_launch": cbytes, "dram_bytes_per_step": scale * (a + b + cbytes),
               "note": "physical DRAM bytes of one bench step = (chunks scaled to the step's configurations) x (cc_fk_kernel + cc_broad_kernel + cc_traverse_kernel)"}
    json.dump(traffic, open(os.path.join(P, "traffic_%s.json" % workload), "w"), indent=1)
    print(json.dumps(phys, indent=1))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
