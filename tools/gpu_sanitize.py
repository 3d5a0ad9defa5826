"""Small end-to-end exercise of every kernel flavour, meant to run under compute-sanitizer (memcheck / racecheck / initcheck).
Run it twice: as is (small batches take the split traversal with its in-warp broadphase) and with CC_TUNE_SPLIT=1 (the full-batch
path at small sizes: cc_fk_kernel / cc_fk2_kernel -> cc_broad_kernel -> cc_traverse_kernel reading candidate bits)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_scene, sample_configs
from openrave_b200 import planning as pl
from openrave_b200 import ravemath as rm
from openrave_b200.checker import CudaCollisionChecker, BATCH_ENV, BATCH_SELF, BATCH_FIRST_HIT

env = load_scene("lab1"); robot = env.GetRobots()[0]
cc = CudaCollisionChecker(env, 0); env.SetCollisionChecker(cc)
q = sample_configs(robot, 3000, seed=1)
for n in (1, 40, 700, 3000):
    r = cc.CheckCollisionBatch(robot, q[:n])
    r = cc.CheckCollisionBatch(robot, q[:n], flags=BATCH_ENV | BATCH_SELF | BATCH_FIRST_HIT)
print("collision ok", int((r["verdict"] != 0).sum()), "tables", r["query"].table_count())
r = cc.CheckCollisionBatch(robot, q[:300], all_pairs=True)
d = cc.DistanceBatch(robot, q[:200])
d2 = cc.DistanceBatch(robot, q[:200], rel_err=200.0, abs_err=0.001)
w = r["query"].within_tolerance(0.02, q[:40])
print("distance ok", float(d["distance"].mean()), float(d2["distance"].mean()), int(w.sum()))
robot.SetDOFValues([0.0, 0.6, 0.0, 1.4, 0.0, 0.3, 0.0, 0.0, 0.0, 0.0, 0.0])
mug = [b for b in env.GetBodies() if b.name == "mug1"][0]
mug.SetTransform(rm.t_mul(robot.GetLink("handbase").t, (rm.IDENT_Q, (0.0, 0.0, 0.10))))
robot.Grab(mug, robot.GetLink("handbase"))
r = cc.CheckCollisionBatch(robot, q[:500])
print("grabbed ok", int((r["verdict"] != 0).sum()), env.CheckCollision(robot), robot.CheckSelfCollision())
robot.ReleaseAllGrabbed()
env2 = load_scene("wamtest1"); robot2 = env2.GetRobots()[0]
robot2.SetActiveDOFs(robot2.GetActiveManipulator().GetArmIndices())
cc2 = CudaCollisionChecker(env2, 0); env2.SetCollisionChecker(cc2)
q2 = sample_configs(robot2, 200, seed=3, active=robot2.GetActiveDOFIndices())
free = q2[cc2.CheckCollisionBatch(robot2, q2)["verdict"] == 0]
res = pl.BirrtPlanner(robot2, cc2).PlanBatch(free[0:8:2], free[1:8:2], seeds=np.arange(4))
print("planner ok", [x["status"] for x in res])

# round 2: PR2 with arm + torso active (two-phase FK kernel, certificates + table lookups in the broadphase kernel), an affine
# base query, the analytic-primitive mode
env3 = load_scene("pr2test1"); robot3 = env3.GetRobots()[0]
arm3 = [m for m in robot3.GetManipulators() if m.GetName() == "rightarm_torso"][0].GetArmIndices()
robot3.SetActiveDOFs(arm3)
cc3 = CudaCollisionChecker(env3, 0); env3.SetCollisionChecker(cc3)
q3 = sample_configs(robot3, 1500, seed=5, active=arm3)
r3 = cc3.CheckCollisionBatch(robot3, q3, active=arm3)
r3b = cc3.CheckCollisionBatch(robot3, q3[:400], active=arm3, all_pairs=True)
print("pr2 ok", int((r3["verdict"] != 0).sum()), "tables", r3["query"].table_count())
from openrave_b200.flatten import DOF_X, DOF_Y, DOF_RotationAxis
robot3.SetActiveDOFs(arm3, DOF_X | DOF_Y | DOF_RotationAxis, [0.0, 0.0, 1.0])
from test_affine_dofs import _sample as sample_affine
q4 = sample_affine(robot3, arm3, DOF_X | DOF_Y | DOF_RotationAxis, 600, seed=7)
r4 = cc3.CheckCollisionBatch(robot3, q4)
print("affine ok", int((r4["verdict"] != 0).sum()))
cc.SetAnalyticPrimitives(True)
r5 = cc.CheckCollisionBatch(robot, q[:1500])
r5b = cc.CheckCollisionBatch(robot, q[:200], all_pairs=True)
print("analytic ok", int((r5["verdict"] != 0).sum()))
