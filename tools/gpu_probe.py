#!/usr/bin/env python
"""Device-side work counters for a workload: box tests / triangle tests per configuration, kernel time."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from openrave_b200.checker import CudaCollisionChecker, BATCH_ENV, BATCH_SELF, BATCH_ALL_PAIRS

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
env, robot, _, seed = bench.load_world("wam_lab1_100k")
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
