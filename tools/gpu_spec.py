import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from openrave_b200 import planning as pl
from openrave_b200.checker import CudaCollisionChecker
scene, nproblems, seed, desc = bench.PLAN_WORKLOADS["wam_wamtest1_birrt100"]
env, robot, cc, q = bench.plan_problems(scene, nproblems, seed, 0, lambda e: CudaCollisionChecker(e, 0))
free = q[cc.CheckCollisionBatch(robot, q)["verdict"] == 0]
starts, goals = free[0:2 * nproblems:2].copy(), free[1:2 * nproblems:2].copy()
seeds = np.arange(nproblems, dtype=np.uint32)
base = None
for spec in (0, 32, 16, -16, -8, -4):
    P = pl.BirrtPlanner(robot, cc, speculate_steps=spec)
    for _ in range(3): res = P.PlanBatch(starts, goals, seeds)
    t0 = time.perf_counter()
    for _ in range(5): res = P.PlanBatch(starts, goals, seeds)
    dt = (time.perf_counter() - t0) / 5
    sig = [r["path"].tobytes() for r in res]
    if base is None: base = sig
    print("speculate_steps %4d: %.0f plans/s, rounds %d, submitted %d, same plans %s" % (spec, nproblems / dt, P.totals["rounds"], P.totals["configs_submitted"], sig == base), flush=True)
