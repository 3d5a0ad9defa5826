"""PR2 / HIRO-NX scenes (COLLADA ingestion, openrave_b200/colladascene.py) pinned against the reference's own tests (not gpu).

Goldens: /root/reference/test/test_planning.py:331-375 (`test_planwithcollision`, data/pr2test1.env.xml):
  * l_shoulder_lift_joint = 0.187                         -> env.CheckCollision(robot) is True        (:339-340)
  * leftarm = [1.34046301, 0.94535038, 3.03934583, -1.30743665, 0, 0, 0]  -> robot.CheckSelfCollision() (:356-357)
  * leftarm = [1.34046301, -0.52360053, 0.03541482, -2.32130534, 0, 0, 0] -> no self collision           (:371-372)
Structure pins: SURVEY.md §8 (PR2: 88 links, 87 joints = 39 DOF + 48 passive, 6 mimic formulas, 31 ignore_link_pair;
HIRO-NX: 24 links, 23 revolute DOF) and the manipulator arm indices openravepy reports for these robots.
The fixtures are produced by tools/make_fixtures.py from the reference's data files.
"""
import numpy as np
import pytest

import oracle
from conftest import have_ref_oracle, load_scene
from openrave_b200.environment import CollisionReport


@pytest.fixture(scope="module")
def pr2env():
    env = load_scene("pr2test1")
    env.SetCollisionChecker(oracle.OracleChecker(env, "reference" if have_ref_oracle() else "port"))
    return env


def test_pr2_structure(pr2env):
    robot = pr2env.GetRobots()[0]
    assert robot.GetName() == "pr2"
    assert len(robot.GetLinks()) == 88
    assert robot.GetDOF() == 39 and len(robot.GetPassiveJoints()) == 48
    assert sum(1 for j in robot.GetPassiveJoints() if j.mimic is not None) == 6
    assert len(robot.forced_adjacent) == 31
    assert sum(l.idx.shape[0] for l in robot.GetLinks()) == 51830
    manips = {m.GetName(): m for m in robot.GetManipulators()}
    assert manips["leftarm"].GetArmIndices() == [15, 16, 17, 18, 19, 20, 21]
    assert manips["rightarm"].GetArmIndices() == [27, 28, 29, 30, 31, 32, 33]
    assert manips["rightarm_torso"].GetArmIndices() == [12, 27, 28, 29, 30, 31, 32, 33]
    tj = robot.GetJoint("torso_lift_joint")
    assert tj.GetDOFIndex() == 12 and abs(tj.lower - 0.0115) < 1e-12 and abs(tj.upper - 0.305) < 1e-12
    # zero pose clamped into the limits at load (kinbody.cpp:5026-5048)
    assert abs(robot.GetDOFValues()[12] - 0.0115) < 1e-12
    assert [b.GetName() for b in pr2env.GetBodies()] == ['pr2', 'floorwalls', 'pole', 'pole2', 'pole3', 'wall1', 'segway', 'mug1',
                                                         'dishwasher_table', 'table']


def test_pr2_collision_goldens(pr2env):
    env = pr2env
    robot = env.GetRobots()[0]
    default = robot.GetDOFValues()
    assert not env.CheckCollision(robot)
    assert not robot.CheckSelfCollision()
    robot.SetDOFValues([0.187], [robot.GetJoint('l_shoulder_lift_joint').GetDOFIndex()])
    rep = CollisionReport()
    assert env.CheckCollision(robot, report=rep)                                       # test_planning.py:339-340
    assert ('l_forearm_link', 'segway') in [(a.name, b.body.name) for a, b in rep.vLinkColliding]
    robot.SetDOFValues(default)
    larm = [m for m in robot.GetManipulators() if m.GetName() == 'leftarm'][0].GetArmIndices()
    robot.SetDOFValues([1.34046301, 0.94535038, 3.03934583, -1.30743665, 0, 0, 0], larm)
    assert robot.CheckSelfCollision()                                                  # :356-357
    robot.SetDOFValues([1.34046301, -0.52360053, 0.03541482, -2.32130534, 0, 0, 0], larm)
    assert not robot.CheckSelfCollision()                                              # :371-372
    robot.SetDOFValues(default)


def test_pr2_mimic_gripper(pr2env):
    robot = pr2env.GetRobots()[0]
    default = robot.GetDOFValues()
    j = robot.GetJoint('l_gripper_l_finger_joint')
    robot.SetDOFValues([0.3], [j.GetDOFIndex()])
    from openrave_b200 import ravemath as rm
    import math
    # l_gripper_r_finger_joint = 0 + l_gripper_l_finger_joint (formula joint70): relative rotation of its two links = 0.3
    mj = robot.GetJoint('l_gripper_r_finger_joint')
    a, b = mj.GetHierarchyParentLink(), mj.GetHierarchyChildLink()
    robot.SetDOFValues([0.0], [j.GetDOFIndex()])
    t0 = rm.t_mul(rm.t_inv(a.t), b.t)
    robot.SetDOFValues([0.3], [j.GetDOFIndex()])
    t1 = rm.t_mul(rm.t_inv(a.t), b.t)
    d = rm.t_mul(rm.t_inv(t0), t1)
    ang = 2 * math.atan2(math.sqrt(sum(x * x for x in d[0][1:])), abs(d[0][0]))
    assert abs(ang - 0.3) < 1e-12
    robot.SetDOFValues(default)


def test_hironx_structure_and_home():
    env = load_scene("hironxtable")
    env.SetCollisionChecker(oracle.OracleChecker(env, "port"))
    robot = env.GetRobots()[0]
    assert robot.GetName() == "HiroNX"
    assert len(robot.GetLinks()) == 24 and robot.GetDOF() == 23 and len(robot.GetPassiveJoints()) == 0
    assert sum(l.idx.shape[0] for l in robot.GetLinks()) == 24116
    manips = {m.GetName(): m for m in robot.GetManipulators()}
    assert manips["rightarm"].GetArmIndices() == [3, 4, 5, 6, 7, 8]
    assert manips["leftarm_torso"].GetArmIndices() == [0, 13, 14, 15, 16, 17, 18]
    assert not env.CheckCollision(robot)          # test_collision.py:26-35 loads this scene and expects a free start
    assert not robot.CheckSelfCollision()
    # regression pin (no external golden exists: the reference cannot run here): 24 links -> 276 pairs, 252 of them non-adjacent and
    # apart at the load pose.  The same list comes back through the CUDA checker's probe (tests/test_gpu_parity_robots.py).
    assert len(robot.GetNonAdjacentLinks(0)) == 252


def test_pr2_nonadjacent_pair_count():
    """901 self pairs: 88 links -> 3828 pairs, minus adjacency (joined links, the 31 ignore_link_pair extras, links one non-static
    joint apart, geometry-less links) and minus the pairs in contact at the load pose (kinbody.cpp:4940-5010, 5532-5612).  Regression
    pin; the round-1 review counted the same number independently."""
    env = load_scene("pr2test1")
    env.SetCollisionChecker(oracle.OracleChecker(env, "port"))
    robot = env.GetRobots()[0]
    na = robot.GetNonAdjacentLinks(0)
    assert len(na) == 901 and len(set(na)) == 901
    assert all((v & 0xffff) < (v >> 16) < 88 for v in na)
