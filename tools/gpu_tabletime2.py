import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_scene, sample_configs
from openrave_b200.checker import CudaCollisionChecker
order = sys.argv[1:] or ["hironxtable", "lab1", "pr2test1"]
manips = {"pr2test1": "rightarm_torso", "lab1": None, "hironxtable": None}
for rep in range(2):
    for scene in order:
        manip = manips[scene]
        env = load_scene(scene); robot = env.GetRobots()[0]
        active = None
        if manip: active = [m for m in robot.GetManipulators() if m.GetName() == manip][0].GetArmIndices(); robot.SetActiveDOFs(active)
        cc = CudaCollisionChecker(env, 0); env.SetCollisionChecker(cc)
        q = sample_configs(robot, 20000, seed=1, active=active)
        bq = cc.bind(robot, active)
        cc.CheckCollisionBatch(robot, q, active=active)
        n = bq.table_count()
        print(rep, scene, "certificates %.1f ms (%d pairs, %d dropped)" % (bq.certificate_build_ms, n, bq.dropped_pairs), flush=True)
