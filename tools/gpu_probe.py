#!/usr/bin/env python
"""Device-side work counters for a workload: box tests / triangle tests per configuration, kernel time."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from openrave_b200.checker import CudaCollisionChecker, BATCH_ENV, BATCH_SELF, BATCH_ALL_PAIRS

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
workload = sys.argv[2] if len(sys.argv) > 2 else "wam_lab1_100k"
env, robot, _, seed = bench.load_world(workload)
print("==", workload, n)
cc = CudaCollisionChecker(env, 0)
env.SetCollisionChecker(cc)
q = bench.sample(robot, n, seed)
bq = cc.bind(robot)
for name, flags in (("env+self", BATCH_ENV | BATCH_SELF), ("env", BATCH_ENV), ("self", BATCH_SELF),
                    ("env+self all-pairs", BATCH_ENV | BATCH_SELF | BATCH_ALL_PAIRS)):
    for counters in (0, 1):
        cc.SetCounters(counters)
        s0 = cc.GetStatistics()
        for _ in range(3):
            v, fe, fs, _ = bq.check(q, flags)
        s1 = cc.GetStatistics()
        print("%-20s counters=%d kernel %.3f ms  bv/config %.1f  tri/config %.2f  colliding %.3f" % (
            name, counters, s1["last_kernel_ms"], (s1["bv_tests"] - s0["bv_tests"]) / 3.0 / n,
            (s1["tri_tests"] - s0["tri_tests"]) / 3.0 / n, (v != 0).mean()))

# per-configuration work distribution (traversal iterations): what the tail of the persistent kernel is made of
bq.trace_work(True)
cc.SetCounters(0)
v, fe, fs, _ = bq.check(q, BATCH_ENV | BATCH_SELF)
w = bq.work_trace(n).astype(np.int64)
bq.trace_work(False)
srt = np.sort(w)[::-1]
print("work trace: total iterations %d, mean %.1f, zero-work configs %.3f" % (w.sum(), w.mean(), (w == 0).mean()))
print("  percentiles 50/90/99/99.9/max:", [int(np.percentile(w, p)) for p in (50, 90, 99, 99.9)], int(w.max()))
print("  top 10:", srt[:10].tolist(), " share of top 0.1%%: %.3f" % (srt[:max(1, n // 1000)].sum() / max(1, w.sum())))
print("  heaviest configs collide?", [(int(i), int(v[i])) for i in np.argsort(-w)[:6]])
for name, sel in (("free", v == 0), ("env only", v == 1), ("self only", v == 2), ("both", v == 3)):
    if sel.any():
        print("  %-10s share %.3f  mean iterations %.1f  share of work %.3f" % (name, sel.mean(), w[sel].mean(), w[sel].sum() / max(1, w.sum())))
