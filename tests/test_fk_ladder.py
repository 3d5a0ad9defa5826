"""Verification tier V1: the FK restatements against closed-form facts, CPU only.

Two independent restatements of KinBody::SetDOFValues (src/libopenrave/kinbody.cpp:2121-2485) exist on the host: the harness
(openrave_b200/kinbody.py, Python) and the oracle (oracle/oracle.cpp, C++).  They must agree with each other to rounding, return
the scene file's link poses at the load values, and a single-joint sweep may move exactly the links that joint affects
(KinBody::DoesAffect; the reference checks the same property in test/test_kinematics.py:263-322)."""
import numpy as np
import pytest

import oracle
from conftest import load_scene, sample_configs
from openrave_b200 import flatten

SCENES = ["lab1", "pr2test1", "hironxtable"]


def _poses(robot):
    return flatten.transforms8(robot.GetLinkTransformations())


def _same_pose(a, b, tol):
    """quaternions up to sign, translations absolutely"""
    dq = np.minimum(np.abs(a[..., :4] - b[..., :4]).max(-1), np.abs(a[..., :4] + b[..., :4]).max(-1))
    return (dq <= tol) & (np.abs(a[..., 4:7] - b[..., 4:7]).max(-1) <= tol)


@pytest.mark.parametrize("scene", SCENES)
def test_load_pose_single_joint_sweeps_and_two_restatements(scene):
    env = load_scene(scene)
    robot = env.GetRobots()[0]
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    ndof = robot.GetDOF()
    q0 = np.array(robot.GetDOFValues(), dtype=np.float64)
    load = _poses(robot)
    alld = list(range(ndof))
    # the load values reproduce the scene file's link poses in both restatements
    fk0 = orc.check_batch(robot, q0[None], active=alld, flags=0, want_fk=True)["fk"][0]
    assert _same_pose(fk0, load, 1e-12).all()
    robot.SetDOFValues(q0.tolist())
    assert _same_pose(_poses(robot), load, 1e-12).all()
    # one joint at a time: exactly the links it affects move
    lo, hi = (np.array(x) for x in robot.GetDOFLimits())
    rng = np.random.default_rng(3)
    moved_any = 0
    for d in rng.permutation(ndof)[:12]:
        q = q0.copy()
        span = hi[d] - lo[d]
        if not span > 1e-6:
            continue
        q[d] = lo[d] + (0.25 if abs(q0[d] - (lo[d] + 0.25 * span)) > 0.1 * span else 0.75) * span
        fk = orc.check_batch(robot, q[None], active=alld, flags=0, want_fk=True)["fk"][0]
        robot.SetDOFValues(q.tolist())
        assert _same_pose(_poses(robot), fk, 1e-12).all()                  # harness == oracle
        affected = np.array([bool(robot.DoesAffect(robot.joints[d].jointindex, l.index)) for l in robot.links])
        same = _same_pose(fk, load, 1e-12)
        assert same[~affected].all()                                      # links the joint does not affect stay put
        moved_any += int((~same[affected]).sum())
    assert moved_any > 0
    robot.SetDOFValues(q0.tolist())
    # random configurations: the two restatements agree to rounding
    q = sample_configs(robot, 200, seed=9)
    fk = orc.check_batch(robot, q, active=alld, flags=0, want_fk=True)["fk"]
    for k in range(0, 200, 20):
        robot.SetDOFValues(q[k].tolist())
        assert _same_pose(_poses(robot), fk[k], 1e-11).all()
    robot.SetDOFValues(q0.tolist())


def _rel(fk, parent, child):
    """pose of `child` in `parent`'s frame, [.., 8] rows {quat wxyz, trans, 0}"""
    from openrave_b200 import ravemath as rm
    tp = (tuple(fk[parent, 0:4]), tuple(fk[parent, 4:7]))
    tc = (tuple(fk[child, 0:4]), tuple(fk[child, 4:7]))
    r = rm.t_mul(rm.t_inv(tp), tc)
    return np.array(list(r[0]) + list(r[1]) + [0.0])


def _overshoot_world():
    """lab1's WAM with the upper limit of the mimic finger joint JF1mimic (value = JF1/3 + 0.8727) pulled to 1e-5 rad BELOW the
    value the equation reaches at JF1's own upper limit."""
    env = load_scene("lab1")
    robot = env.GetRobots()[0]
    jm = [j for j in robot.sorted_joints_all if j.name == "JF1mimic"][0]
    src = robot.GetJoint("JF1")
    a, b = jm.mimic.affine()[1:]
    jm.upper = a * src.upper + b - 1e-5
    return env, robot, jm, src, a, b


def test_mimic_value_beyond_its_limit_is_clamped_like_the_reference():
    """kinbody.cpp:2326-2375 with the double build's tolerances (g_fEpsilonJointLimit = 1e-12, g_fEpsilonEvalJointLimit = 1.8e-10,
    libopenrave.h:141-142): a mimic value 1e-5 rad past its limit is CLAMPED to the limit (it was left alone while the constants
    read 1e-7 / 5e-4), one 1e-11 past it is kept (the reference's sliver between the two tolerances), one 1e-13 past it snaps."""
    env, robot, jm, src, a, b = _overshoot_world()
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    alld = list(range(robot.GetDOF()))
    q0 = np.array(robot.GetDOFValues(), dtype=np.float64)
    parent, child = jm.links[0].index, jm.links[1].index

    def poses(jf1):
        q = q0.copy()
        q[src.dofindex] = jf1
        fk = orc.check_batch(robot, q[None], active=alld, flags=0, want_fk=True)["fk"][0]
        robot.SetDOFValues(q.tolist())
        assert _same_pose(_poses(robot), fk, 1e-12).all()                  # harness == oracle, whatever the rule does
        return _rel(fk, parent, child)

    at_limit = poses((jm.upper - b) / a)                                   # the equation lands exactly on the limit
    over = poses(src.upper)                                                # 1e-5 rad beyond it
    assert _same_pose(over, at_limit, 1e-14)
    sliver = poses((jm.upper + 1e-11 - b) / a)
    assert not _same_pose(sliver, at_limit, 1e-13)                         # kept as evaluated (off by ~5e-12)
    snap = poses((jm.upper + 2e-13 - b) / a)
    assert _same_pose(snap, at_limit, 1e-14)
    robot.SetDOFValues(q0.tolist())


@pytest.mark.gpu
def test_gpu_mimic_value_beyond_its_limit():
    from openrave_b200.checker import CudaCollisionChecker
    env, robot, jm, src, a, b = _overshoot_world()
    cc = CudaCollisionChecker(env, 0)
    orc = oracle.OracleChecker(env, "port")
    orc.InitEnvironment()
    alld = list(range(robot.GetDOF()))
    q0 = np.array(robot.GetDOFValues(), dtype=np.float64)
    vals = [src.upper, (jm.upper - b) / a, (jm.upper + 1e-11 - b) / a, (jm.upper + 2e-13 - b) / a, src.upper - 3e-6]
    q = np.repeat(q0[None], len(vals), axis=0)
    q[:, src.dofindex] = vals
    got = cc.ForwardKinematicsBatch(robot, q, active=alld)
    ref = orc.check_batch(robot, q, active=alld, flags=0, want_fk=True)["fk"]
    assert _same_pose(got, ref, 1e-12).all()
    parent, child = jm.links[0].index, jm.links[1].index
    rel = [_rel(got[k], parent, child) for k in range(len(vals))]
    assert _same_pose(rel[0], rel[1], 1e-14) and not _same_pose(rel[2], rel[1], 1e-13) and _same_pose(rel[3], rel[1], 1e-14)
