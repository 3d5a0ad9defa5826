"""The oracle pinned against the reference's own known-answer tests for lab1 (not gpu).

Goldens: /root/reference/test/test_collision.py:183-188 (known collision), :190-205 (home pose free, mug1 in
collision), :223-251 (pqp distances 0.38737 / 0.0027169 / 0.29193971893003506 and the nearest link pairs).
The distance values need the unmodified PQP build (oracle/_ref) because pqprave runs PQP_Distance with
rel_err=200, abs_err=0.001 (collisionPQP.h:48-49) — an approximate search that depends on the BVH.
"""
import numpy as np
import pytest

import oracle
from conftest import have_ref_oracle
from openrave_b200.environment import CollisionReport

KNOWN_COLLISION = [-8.44575603e-02, 1.48528347e+00, -5.09108824e-08, 6.48108822e-01, -4.57571203e-09, -1.04008750e-08,
                   7.26855048e-10, 5.50807826e-08, 5.50807826e-08, -1.90689327e-08, 0.00000000e+00]


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_lab1_verdict_goldens(lab1, kind):
    if kind == "reference" and not have_ref_oracle():
        pytest.skip("oracle/_ref not built")
    env = lab1
    env.SetCollisionChecker(oracle.OracleChecker(env, kind))
    robot = env.GetRobots()[0]
    assert robot.GetDOF() == 11 and len(robot.GetLinks()) == 18
    assert not env.CheckCollision(robot)                       # test_collision.py:203
    assert env.CheckCollision(env.GetKinBody('mug1'))          # test_collision.py:206
    assert not robot.CheckSelfCollision()
    robot.SetDOFValues(KNOWN_COLLISION)
    report = CollisionReport()
    assert env.CheckCollision(robot, report=report)            # test_collision.py:183-188
    assert len(report.vLinkColliding) >= 1
    # determinism (test_collision.py:37-68): same query twice, same verdict and pairs
    report2 = CollisionReport()
    assert env.CheckCollision(robot, report=report2)
    assert [(a.name, b.body.name) for a, b in report.vLinkColliding] == [(a.name, b.body.name) for a, b in report2.vLinkColliding]


def test_lab1_nonadjacent_pairs(lab1):
    env = lab1
    env.SetCollisionChecker(oracle.OracleChecker(env, "port"))
    robot = env.GetRobots()[0]
    pairs = robot.GetNonAdjacentLinks(0)
    assert len(pairs) == 122
    # CompareNonAdjacentFarthest (kinbody.cpp:5512-5530): farthest link indices first
    assert (pairs[0] & 0xffff, pairs[0] >> 16) == (0, 17)
    keys = [(-((p >> 16) - (p & 0xffff)), -(p >> 16), -(p & 0xffff)) for p in pairs]
    assert keys == sorted(keys)
    for p in pairs:
        assert not robot.AreAdjacentLinks(p & 0xffff, p >> 16)


@pytest.mark.skipif(not have_ref_oracle(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_lab1_pqp_distance_goldens(lab1):
    env = lab1
    cc = oracle.OracleChecker(env, "reference")
    env.SetCollisionChecker(cc)
    robot = env.GetRobots()[0]
    manip = robot.GetActiveManipulator()
    ee = manip.GetEndEffector()
    others = [(b, l) for b in env.GetBodies() if b is not robot for l in b.links]
    d, bname = min((cc.ref_distance(ee, l), b.name) for b, l in others)
    assert abs(d - 0.38737) < 0.01 and bname == 'pole'                       # test_collision.py:233-236
    d, lname, bname = min((cc.ref_distance(rl, l), rl.name, b.name) for rl in robot.links for b, l in others)
    assert abs(d - 0.0027169) < 0.01 and (lname, bname) == ('segway', 'floorwalls')   # :238-242
    arm = manip.GetArmIndices()
    assert arm == [0, 1, 2, 3, 4, 5, 6]
    active = [l for l in robot.links if any(robot.DoesAffect(dd, l.index) for dd in arm)]
    d, lname, bname = min((cc.ref_distance(rl, l), rl.name, b.name) for rl in active for b, l in others)
    assert abs(d - 0.29193971893003506) < 1e-9 and (lname, bname) == ('wam1', 'pole')  # :244-250 (golden to 17 digits)


def test_mimic_fingers_follow_equation(lab1):
    """barretthand.kinbody.xml: JFxmimic = JFx/3 + 0.8727; test/test_kinematics.py:835-874 style check."""
    robot = lab1.GetRobots()[0]
    import math
    from openrave_b200 import ravemath as rm
    q = [0.0] * 11
    q[7] = 1.2
    robot.SetDOFValues(q)
    f01, f02 = robot.GetLink('Finger0-1'), robot.GetLink('Finger0-2')
    trel = rm.t_mul(rm.t_inv(f01.t), f02.t)
    ang = 2 * math.atan2(math.sqrt(sum(x * x for x in trel[0][1:])), trel[0][0])
    assert abs(ang - (1.2 / 3 + 0.8727)) < 1e-12
