import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_scene, sample_configs
from openrave_b200.checker import CudaCollisionChecker
for scene, manip in (("pr2test1", "rightarm_torso"), ("lab1", None), ("hironxtable", None)):
    env = load_scene(scene); robot = env.GetRobots()[0]
    active = None
    if manip: active = [m for m in robot.GetManipulators() if m.GetName() == manip][0].GetArmIndices(); robot.SetActiveDOFs(active)
    cc = CudaCollisionChecker(env, 0); env.SetCollisionChecker(cc)
    q = sample_configs(robot, 20000, seed=1, active=active)
    t0 = time.perf_counter(); bq = cc.bind(robot, active); t1 = time.perf_counter()
    cc.CheckCollisionBatch(robot, q[:100], active=active); t2 = time.perf_counter()
    cc.CheckCollisionBatch(robot, q, active=active); t3 = time.perf_counter()
    cc.CheckCollisionBatch(robot, q, active=active); t4 = time.perf_counter()
    n = bq.table_count()
    print(scene, "bind %.1f ms, first small batch %.1f ms, batch that builds the tables %.1f ms, same batch after %.1f ms, tables+dropped %d (dropped %d)" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), n, bq.dropped_pairs))
