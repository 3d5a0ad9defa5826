#!/usr/bin/env python
"""Quick start on a B200: the calls an OpenRAVE user makes, against the ctypes twin of the plugin.

  python examples/quickstart.py            # needs a CUDA device (the product has no CPU path)

Inside a real OpenRAVE the same sequence goes through plugins/cudacollision (RaveCreateCollisionChecker(env, "cudacollision"),
see INTEGRATION.md); here the scene comes from the committed fixtures, parsed by our own loaders.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from openrave_b200 import planning as pl                                    # noqa: E402
from openrave_b200.checker import RaveCreateCollisionChecker                # noqa: E402
from openrave_b200.environment import CollisionReport, Environment          # noqa: E402


def main():
    env = Environment()
    env.Load(os.path.join(ROOT, "tests", "golden", "scenes", "wamtest1.npz"))       # or an OpenRAVE env.xml / .zae with $OPENRAVE_DATA
    cc = RaveCreateCollisionChecker(env, "cudacollision 0")
    env.SetCollisionChecker(cc)                                              # InitEnvironment: meshes -> BVHs -> HBM
    robot = env.GetRobots()[0]
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())

    # single-state calls, unchanged from the reference API
    report = CollisionReport()
    print("robot in collision:", env.CheckCollision(robot, report=report), " self:", robot.CheckSelfCollision(report))

    # the batched entry: N active-DOF vectors at once -> verdict bits + first colliding pairs in the reference's order
    lo, hi = robot.GetDOFLimits()
    act = robot.GetActiveDOFIndices()
    lo, hi = np.array([lo[i] for i in act]), np.array([hi[i] for i in act])
    q = lo + np.random.default_rng(0).random((200000, len(act))) * (hi - lo)
    res = cc.CheckCollisionBatch(robot, q)
    v = res["verdict"]
    print("%d configurations: %.1f %% env collisions, %.1f %% self collisions" % (len(q), 100 * (v & 1).mean(), 100 * ((v >> 1) & 1).mean()))

    # CO_Distance for a batch: exact minimum separation, the link pair that attains it, the closest points
    d = cc.DistanceBatch(robot, q[:20000])
    k = int(np.argmax(d["distance"]))
    print("clearance: median %.4f m, largest %.4f m (link %d against body %d link %d)" % (
        float(np.median(d["distance"])), float(d["distance"][k]), *[int(x) for x in d["pair"][k]]))

    # several GPUs of one box behind the same object: RaveCreateCollisionChecker(env, "cudacollision 0,1,2,3") — every call
    # above stays the same, batches are range-sharded inside the library

    # an edge check (CheckPathAllConstraints) and a plan, both on the batched entry
    free = q[v == 0]
    planner = pl.BirrtPlanner(robot, cc)
    print("edge:", planner.CheckEdges(free[0], free[1], pl.IT_Closed)[0])
    plan = planner.PlanPath(free[0], free[1], seed=1)
    print("plan: status %d, %d waypoints, length %.3f (raw %.3f), %d configurations checked" % (
        plan["status"], plan["waypoints"], plan["length"], plan["length_raw"], plan["configs_submitted"]))


if __name__ == "__main__":
    main()
