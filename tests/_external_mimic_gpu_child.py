"""Child process of tests/test_external_mimic.py::test_gpu_equations_evaluated_by_the_library (needs a CUDA device)."""
import ctypes

import numpy as np

from openrave_b200 import planning as pl
from openrave_b200.checker import CudaCollisionChecker
from test_external_mimic import _configs, _nonlinear_world


def main():
    env, robot = _nonlinear_world()
    arm = list(robot.GetActiveManipulator().GetArmIndices()) + [robot.GetJoint("JF1").dofindex, robot.GetJoint("JF2").dofindex]
    robot.SetActiveDOFs(arm)
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    assert cc.RegisterMimicEquations(robot, arm) == 2
    bq = cc.bind(robot, arm)
    assert bq.info.nexternal == 2 and bq.info.ndof_active == len(arm) + 2
    q = np.ascontiguousarray(_configs(robot, 2000, 11)[:, arm])
    rows = bq.rows(q)
    out = np.zeros_like(rows)
    assert cc.lib.cc_query_expand_rows(bq.handle, q.ctypes.data_as(ctypes.c_void_p), len(q), out.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.abs(out - rows).max() < 1e-14
    planner = pl.BirrtPlanner(robot, cc)
    twin = pl.DynamicsCollisionConstraint(robot, cc)
    q0 = q[0:400:2]
    q1 = q0 + 0.2 * (q[1:400:2] - q0)
    res = planner.CheckEdges(q0, q1, pl.IT_Closed, options=0xffff, perturbation=0.0)
    want = twin.CheckMany(list(zip(q0, q1)), pl.IT_Closed, options=0xffff)
    assert [r["return_code"] for r in res] == [int(w) & 0xffffffff for w in want]
    assert 0 < sum(1 for w in want if w) < len(want)
    print("EXTERNAL_MIMIC_LIBRARY_PATH_OK")


if __name__ == "__main__":
    main()
