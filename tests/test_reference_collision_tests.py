"""The reference's own collision tests (test/test_collision.py) replayed through the oracle (CPU) and through the CUDA checker
(GPU) — SURVEY.md §8c lists them as the known answers of this path:
  :70-83   test_selfcollision            mug2 moved onto mug1; Grab does not create self-collisions
  :85-91   test_attachedbodiescollision  CheckCollision(wam5, wam6): adjacent links do intersect at the load pose
  :190-221 test_collisioncallbacks       a callback sees every reported collision and can veto it (CollisionAction.Ignore)
  :254-272 test_multiplecontacts         CO_AllLinkCollisions: 8 pair infos for the robot, 4 for the end effector
The reference runs them with ode / bullet / fcl; the link-pair counts below are reproduced with pqprave's triangle-level
semantics, which is what the oracle and the CUDA path implement."""
import numpy as np
import pytest

import oracle
from conftest import load_scene
from openrave_b200.environment import CollisionAction, CollisionReport


def _oracle_env(scene):
    env = load_scene(scene)
    env.SetCollisionChecker(oracle.OracleChecker(env, "port"))
    return env


def _cuda_env(scene):
    from openrave_b200.checker import CudaCollisionChecker
    env = load_scene(scene)
    env.SetCollisionChecker(CudaCollisionChecker(env, 0))
    return env


FACTORIES = [pytest.param(_oracle_env, id="oracle"), pytest.param(_cuda_env, id="cuda", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("make", FACTORIES)
def test_selfcollision(make):
    env = make("lab1")
    target1, target2 = env.GetKinBody("mug1"), env.GetKinBody("mug2")
    target2.SetTransform(target1.GetTransform())
    assert env.CheckCollision(target1)
    robot = env.GetRobots()[0]
    report = CollisionReport()
    assert not robot.CheckSelfCollision(report)
    robot.Grab(target1, robot.GetActiveManipulator().GetEndEffector())
    assert not robot.CheckSelfCollision(report)
    assert not target1.CheckSelfCollision()
    assert env.CheckCollision(target1, report=report)
    assert report.plink1.body is target1 and any(b.body is target2 for _, b in report.vLinkColliding)


@pytest.mark.parametrize("make", FACTORIES)
def test_attachedbodiescollision(make):
    env = make("lab1")
    robot = env.GetRobot("BarrettWAM")
    assert env.CheckCollision(robot.GetLink("wam5"), robot.GetLink("wam6"))


@pytest.mark.parametrize("make", FACTORIES)
def test_collisioncallbacks(make):
    env = make("lab1")
    robot = env.GetRobots()[0]
    mug1 = env.GetKinBody("mug1")
    reports = []

    def collisioncallback(report, fromphysics):
        assert not fromphysics
        reports.append(report)
        return CollisionAction.DefaultAction

    handle = env.RegisterCollisionCallback(collisioncallback)
    assert not env.CheckCollision(robot)
    assert len(reports) == 0
    assert env.CheckCollision(mug1)                    # mug1 rests on the table
    assert len(reports) == 1
    assert reports[0].plink1.body is mug1

    handle2 = env.RegisterCollisionCallback(lambda report, fromphysics: CollisionAction.Ignore)
    assert not env.CheckCollision(mug1)
    del reports[:]
    handle2.Close()
    assert env.CheckCollision(mug1)
    assert len(reports) == 1
    handle.Close()
    assert env.CheckCollision(mug1)
    assert len(reports) == 1
    # CO_IgnoreCallbacks switches them off (collisionPQP.h:569)
    handle3 = env.RegisterCollisionCallback(lambda report, fromphysics: CollisionAction.Ignore)
    cc = env.GetCollisionChecker()
    assert not env.CheckCollision(mug1)
    cc.SetCollisionOptions(cc.GetCollisionOptions() | 0x100)
    assert env.CheckCollision(mug1)
    handle3.Close()


@pytest.mark.parametrize("make", FACTORIES)
def test_multiplecontacts(make):
    env = make("lab1")
    env.GetCollisionChecker().SetCollisionOptions(0x20)           # CollisionOptions.AllLinkCollisions
    robot = env.GetRobots()[0]
    manip = robot.GetManipulators()[0]
    body1, body2 = env.GetKinBody("mug1"), env.GetKinBody("mug2")
    body1.SetTransform(manip.GetEndEffector().GetTransform())
    body2.SetTransform(manip.GetEndEffector().GetTransform())
    report = CollisionReport()
    env.CheckCollision(robot, report=report)
    assert len(report.collisionInfos) == 8
    manip.CheckEndEffectorCollision(report)
    assert len(report.collisionInfos) == 4


@pytest.mark.parametrize("make", FACTORIES)
def test_selfcollision_joinxml(make):
    """:94-181 — a segway box model joined to the WAM through a dummy joint (robot built from an XML string; the fixture is
    that string parsed by our loader, tools/make_fixtures.py): DOFs 1 and 3 at [-0.91, 2.05] fold the arm into the segway."""
    env = make("joinxml_segwaywam")
    robot = env.GetRobots()[0]
    assert [l.name for l in robot.links][:2] == ["segway", "wam0"] and len(robot.links) == 18
    assert not robot.CheckSelfCollision()
    robot.SetDOFValues([-0.91, 2.05], [1, 3])
    report = CollisionReport()
    assert robot.CheckSelfCollision(report)
    assert report.plink1.body is robot and report.plink2.body is robot


def test_excluded_bodies_and_links_oracle():
    """The (body | link, vbodyexcluded, vlinkexcluded) overloads (collisionchecker.h:249-263; pqprave collisionPQP.h:432-498):
    excluding what a body collides with frees it, excluding something else does not, and excluded links are skipped on
    either side.  (CPU, oracle; the CUDA checker shares `excluding_pairs` and its explicit-pair query.)"""
    env = _oracle_env("lab1")
    mug1, mug2, table = env.GetKinBody("mug1"), env.GetKinBody("mug2"), env.GetKinBody("table")
    robot = env.GetRobots()[0]
    report = CollisionReport()
    assert env.CheckCollision(mug1, report=report)                            # mug1 rests on the table
    others = set(b.body for _, b in report.vLinkColliding)
    assert table in others or len(others) >= 1
    assert not env.CheckCollision(mug1, bodyexcluded=list(others), linkexcluded=[])
    assert env.CheckCollision(mug1, bodyexcluded=[robot], linkexcluded=[])
    assert not env.CheckCollision(mug1, bodyexcluded=[], linkexcluded=[l for b in others for l in b.links])
    assert not env.CheckCollision(mug1, bodyexcluded=[], linkexcluded=list(mug1.links))     # its own link excluded: nothing left
    # the link overload, with a report: only the colliding, non-excluded bodies are named
    mug2.SetTransform(mug1.GetTransform())
    rep = CollisionReport()
    assert env.CheckCollision(mug1.links[0], bodyexcluded=list(others), linkexcluded=[], report=rep)
    assert set(b.body for _, b in rep.vLinkColliding) == {mug2}
    assert not env.CheckCollision(mug1.links[0], bodyexcluded=list(others) + [mug2], linkexcluded=[])
    # same answers as the plain overloads when nothing is excluded
    assert env.CheckCollision(robot, bodyexcluded=[], linkexcluded=[]) == env.CheckCollision(robot)
