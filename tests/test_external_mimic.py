"""Mimic joints whose equation is not a * q[dof] + b (kinbody.cpp:2308-2386 evaluates any fparser expression over any joints).

The device's chain program stores affine single-source equations only; every other mimic joint becomes a CC_MIMIC_EXTERNAL joint
whose raw equation value the caller supplies as an extra column of every configuration row (include/cudacollision.h, cc_joint).
The ctypes twin computes those columns (BoundQuery.rows); the library applies the mimic limit rules and everything downstream.
Ground truth here is the harness' SetDOFValues restatement, which evaluates the parsed expression as written, and the oracle in
pose mode on the link poses it yields.
"""
import numpy as np
import pytest

import oracle
from conftest import load_scene
from openrave_b200 import flatten
from openrave_b200.kinbody import MimicExpr, external_mimic_values


def _nonlinear_world():
    """lab1's WAM with two finger mimic joints rewritten: JF1mimic quadratic in JF1 plus a term in JF2 (two sources, non-affine),
    JF2mimic = JF1mimic / 2 + JF3 / 5 (reads a PASSIVE joint that is itself external), JF3mimic left affine."""
    env = load_scene("lab1")
    robot = env.GetRobots()[0]
    byname = {j.name: j for j in robot.all_joints()}
    byname["JF1mimic"].mimic = MimicExpr("JF1*JF1/5 + JF2/7 + 0.3")
    byname["JF2mimic"].mimic = MimicExpr("JF1mimic/2 + JF3/5 + 0.1")
    robot.compute_internal_information()
    return env, robot


def _configs(robot, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = (np.array(x) for x in robot.GetDOFLimits())
    q = lo + (hi - lo) * rng.random((n, robot.GetDOF()))
    q[::7] = lo + (hi - lo) * np.round(rng.random((len(q[::7]), robot.GetDOF())))      # corners: mimic values past their limits
    return q


def _harness_poses(robot, q):
    out = []
    for row in q:
        robot.SetDOFValues(row.tolist())
        out.append(flatten.transforms8(robot.GetLinkTransformations()))
    return np.array(out)


def test_joint_table_marks_what_the_device_cannot_evaluate():
    env, robot = _nonlinear_world()
    with pytest.raises(NotImplementedError):
        flatten.joint_table(robot)
    ext = []
    arr, n = flatten.joint_table(robot, externals=ext)
    assert [j.name for j in ext] == ["JF1mimic", "JF2mimic"]                # dependency order: JF2mimic reads JF1mimic
    rows = {robot.sorted_joints_all[k].name: arr[k] for k in range(n)}
    assert rows["JF1mimic"].mimic_dof == flatten.CC_MIMIC_EXTERNAL and rows["JF1mimic"].reserved == 0 and rows["JF1mimic"].dof_index == -1
    assert rows["JF2mimic"].mimic_dof == flatten.CC_MIMIC_EXTERNAL and rows["JF2mimic"].reserved == 1
    assert rows["JF3mimic"].mimic_dof == robot.GetJoint("JF3").dofindex     # the affine one stays on the device
    # both new sources reach the affects table (kinbody.cpp:4787-4818 follows mimic sources through passive joints)
    aff = np.array(robot.joints_affecting_links)
    link = [j for j in robot.all_joints() if j.name == "JF2mimic"][0].links[1].index
    for name in ("JF1", "JF2", "JF3"):
        assert aff[robot.GetJoint(name).dofindex, link] == 1


def test_external_values_follow_the_scalar_joint_loop():
    env, robot = _nonlinear_world()
    ext = []
    flatten.joint_table(robot, externals=ext)
    q = _configs(robot, 200, 3)
    got = external_mimic_values(robot, q, ext)
    j1, j2 = ext
    d = {n: robot.GetJoint(n).dofindex for n in ("JF1", "JF2", "JF3")}
    for k in range(len(q)):
        e1 = j1.mimic.eval({"JF1": q[k, d["JF1"]], "JF2": q[k, d["JF2"]]})
        v1 = min(max(e1, j1.lower), j1.upper)                               # no value here lands in the 1e-12..1.8e-10 sliver
        e2 = j2.mimic.eval({"JF1mimic": v1, "JF3": q[k, d["JF3"]]})
        assert got[k, 0] == e1 and got[k, 1] == e2
    assert (got[:, 0] > j1.upper).any() and (got[:, 0] < j1.upper).any()    # both sides of the limit are exercised


@pytest.mark.gpu
def test_gpu_external_mimic_fk_and_verdicts():
    from openrave_b200.checker import CudaCollisionChecker, BATCH_ENV, BATCH_SELF
    env, robot = _nonlinear_world()
    q0 = np.array(robot.GetDOFValues(), dtype=np.float64)
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    orc = oracle.OracleChecker(env, "port")
    orc.InitEnvironment()
    alld = list(range(robot.GetDOF()))
    q = _configs(robot, 3000, 5)
    ref_T = _harness_poses(robot, q)
    robot.SetDOFValues(q0.tolist())
    got_T = cc.ForwardKinematicsBatch(robot, q, active=alld)
    assert np.abs(got_T[:, :len(robot.links)] - ref_T).max() < 1e-9
    bq = cc.bind(robot, alld)
    assert bq.info.ndof_active == robot.GetDOF() + 2 and len(bq.externals) == 2
    got = cc.CheckCollisionBatch(robot, q, BATCH_ENV | BATCH_SELF, active=alld)
    ref = orc.check_batch(robot, None, active=alld, flags=BATCH_ENV | BATCH_SELF, want_band=True, poses=ref_T)
    clear = ref["near"] == 0
    assert clear.sum() > 0.9 * len(q)
    assert (got["verdict"][clear] == ref["verdict"][clear]).all()
    assert 0.02 < (ref["verdict"] != 0).mean() < 0.98
    # pose mode on the same poses is the same computation
    v2 = bq.check_poses(ref_T)[0]
    assert (v2[clear] == ref["verdict"][clear]).all()
    # a subset of the DOFs active: the fingers' sources stay at the body's current values
    arm = [robot.GetJoint("JF1").dofindex, 0, 1, 3]
    qa = q[:500][:, arm]
    full = np.repeat(q0[None], len(qa), axis=0)
    full[:, arm] = qa
    ref_Ta = _harness_poses(robot, full)
    robot.SetDOFValues(q0.tolist())
    got_Ta = cc.ForwardKinematicsBatch(robot, qa, active=arm)
    assert np.abs(got_Ta[:, :len(robot.links)] - ref_Ta).max() < 1e-9
    gota = cc.CheckCollisionBatch(robot, qa, BATCH_ENV | BATCH_SELF, active=arm)
    refa = orc.check_batch(robot, None, active=arm, flags=BATCH_ENV | BATCH_SELF, want_band=True, poses=ref_Ta)
    ca = refa["near"] == 0
    assert (gota["verdict"][ca] == refa["verdict"][ca]).all()
    # the oracle's own FK refuses such a body instead of guessing
    with pytest.raises(NotImplementedError):
        orc.check_batch(robot, q[:4], active=alld)


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="written after this round's GPU time was spent: its first run on a device is the driver's "
                                        "(the host code it reaches is pinned device-free in tests/test_mimic_expr.py)")
def test_gpu_equations_evaluated_by_the_library():
    """cc_body_set_mimic_equation + cc_query_expand_rows (what the OpenRAVE shim and the planner host use) against the twin's own
    evaluation on the same device checker: identical rows, and edges checked by the planner host (which expands the rows itself)
    get the return codes the twin's per-sample path gives them.  Runs in a process of its own (tests/_external_mimic_gpu_child.py):
    code that has never met a device does not get to take the rest of the suite down with it."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.dirname(here), here, os.environ.get("PYTHONPATH", "")]))
    r = subprocess.run([sys.executable, os.path.join(here, "_external_mimic_gpu_child.py")], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "EXTERNAL_MIMIC_LIBRARY_PATH_OK" in r.stdout, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
