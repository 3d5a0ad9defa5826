#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of an ncu report (source page), grouped into line ranges.

  python tools/ncu_lines.py rep.ncu-rep [lo-hi:name ...]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    groups = []
    for a in sys.argv[2:]:
        rng, name = a.split(":")
        lo, hi = rng.split("-")
        groups.append((int(lo), int(hi), name))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = [r for r in rows if r and r[0] == "Line No"][0]
    ii = hdr.index("Instructions Executed")
    si = hdr.index("# Samples")
    per = {}
    other_i = 0
    mine = False
    for r in rows:
        if r and r[0] == "File Path":
            mine = r[1].endswith("cc_device.cuh")
            continue
        if len(r) > ii and r[0].isdigit():
            try:
                v = (int(r[ii] or 0), int(r[si] or 0), r[1])
            except ValueError:
                continue
            if mine:
                per[int(r[0])] = v
            else:
                other_i += v[0]
    print("instructions attributed to other files: %d" % other_i)
    tot_i = sum(v[0] for v in per.values())
    tot_s = sum(v[1] for v in per.values())
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    if groups:
        for lo, hi, name in groups:
            gi = sum(v[0] for k, v in per.items() if lo <= k <= hi)
            gs = sum(v[1] for k, v in per.items() if lo <= k <= hi)
            print("%-28s lines %4d-%4d  instr %5.1f%%  samples %5.1f%%" % (name, lo, hi, 100.0 * gi / tot_i, 100.0 * gs / max(1, tot_s)))
    else:
        for k in sorted(per):
            v = per[k]
            if v[0] * 200 > tot_i or v[1] * 200 > tot_s:
                print("%4d instr %5.2f%% samples %5.2f%%  %s" % (k, 100.0 * v[0] / tot_i, 100.0 * v[1] / max(1, tot_s), v[2][:110]))


if __name__ == "__main__":
    main()
