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
