#!/usr/bin/env python
"""Small-batch latency of the host-buffer entry (what a single planner's CheckPathAllConstraints call pays)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openrave_b200.checker import CudaCollisionChecker, PinnedArray, BATCH_ENV, BATCH_SELF
from openrave_b200.environment import Environment

env = Environment(); env.Load(os.path.join(ROOT, "tests", "golden", "scenes", "wamtest1.npz"))
robot = env.GetRobots()[0]
robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
cc = CudaCollisionChecker(env, 0); env.SetCollisionChecker(cc)
bq = cc.bind(robot)
lo, hi = robot.GetActiveDOFLimits(); lo, hi = np.array(lo), np.array(hi)
rng = np.random.default_rng(0)
out = {}
for n in (1, 8, 32, 128, 512, 2048, 8192, 32768):
    q = lo + rng.random((n, len(lo))) * (hi - lo)
    for _ in range(20): bq.check(q)
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps): bq.check(q)
    dt = (time.perf_counter() - t0) / reps
    pv, pe, ps = PinnedArray(n, np.uint32), PinnedArray(n, np.int32), PinnedArray(n, np.int32)
    pq = PinnedArray((n, len(lo)), np.float64); pq.array[:] = q
    for _ in range(20): bq.check(pq.array, out=(pv.array, pe.array, ps.array))
    t0 = time.perf_counter()
    for _ in range(reps): bq.check(pq.array, out=(pv.array, pe.array, ps.array))
    dtp = (time.perf_counter() - t0) / reps
    out[n] = dict(us_pageable=dt * 1e6, us_pinned=dtp * 1e6, kernel_ms=cc.GetStatistics()["last_kernel_ms"])
    print(n, out[n], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "latency.json"), "w"), indent=1)
