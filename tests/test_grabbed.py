"""Grabbed bodies: KinBody::Grab + KinBody::CheckSelfCollision (src/libopenrave/kinbodygrab.cpp,
kinbodycollision.cpp:83-205) and pqprave's walk over every attached body (plugins/pqprave/collisionPQP.h:314-329).

CPU part: the oracle's batched loop (FK + _UpdateGrabbedBodies + the reference's grabbed loops in C++) agrees with the literal
per-call path of the harness (SetDOFValues, then Environment::CheckCollision(robot) and the Python restatement of
KinBody::CheckSelfCollision issuing link-link queries).  GPU part: the CUDA batched entry against the oracle's."""
import numpy as np
import pytest

import oracle
from conftest import load_scene, sample_configs
from openrave_b200 import ravemath as rm


def _grab_world(checker_factory):
    env = load_scene("lab1")
    robot = env.GetRobots()[0]
    cc = checker_factory(env)
    env.SetCollisionChecker(cc)
    home = [0.0, 0.6, 0.0, 1.4, 0.0, 0.3, 0.0, 0.0, 0.0, 0.0, 0.0]
    robot.SetDOFValues(home)
    by = {b.name: b for b in env.GetBodies()}
    hand, fore = robot.GetLink("handbase"), robot.GetLink("wam4")
    # mug1 into the hand (it touches the palm and the fingers: those links must drop out of the non-colliding list),
    # mug2 strapped to the forearm a little off its surface
    by["mug1"].SetTransform(rm.t_mul(hand.t, (rm.IDENT_Q, (0.0, 0.0, 0.10))))
    by["mug2"].SetTransform(rm.t_mul(fore.t, (rm.IDENT_Q, (0.0, 0.22, 0.05))))
    return env, robot, cc, by, hand, fore


def _literal(env, robot, q):
    env_hit, self_hit = [], []
    for row in q:
        robot.SetDOFValues(row.tolist())
        env_hit.append(env.CheckCollision(robot))
        self_hit.append(robot.CheckSelfCollision())
    return np.array(env_hit), np.array(self_hit)


def test_grab_lists_and_follow():
    env, robot, orc, by, hand, fore = _grab_world(lambda e: oracle.OracleChecker(e, "port"))
    assert robot.Grab(by["mug1"], hand)
    g = robot.grabbed[0]
    names = [l.name for l in g.noncolliding]
    assert "handbase" not in names and "wam7" not in names          # rigidly attached to the grabbing link
    assert "wam1" in names and "wam3" in names
    assert robot.IsGrabbing(by["mug1"]) and by["mug1"].IsAttached(robot) and robot.IsAttached(by["mug1"])
    rel = rm.t_mul(rm.t_inv(hand.t), by["mug1"].links[0].t)
    robot.SetDOFValues([0.4, 0.9, -0.2, 1.0, 0.3, -0.2, 0.1, 0.2, 0.2, 0.2, 0.0])
    now = rm.t_mul(rm.t_inv(hand.t), by["mug1"].links[0].t)
    assert np.allclose(rel[0], now[0], atol=1e-12) and np.allclose(rel[1], now[1], atol=1e-12)
    robot.Release(by["mug1"])
    assert not robot.IsGrabbing(by["mug1"]) and not by["mug1"].IsAttached(robot)


def test_oracle_batch_equals_literal_loop_with_grabbed_bodies():
    env, robot, orc, by, hand, fore = _grab_world(lambda e: oracle.OracleChecker(e, "port"))
    q = sample_configs(robot, 160, seed=17)
    base = orc.check_batch(robot, q)["verdict"]
    robot.Grab(by["mug1"], hand)
    robot.Grab(by["mug2"], fore)
    assert any(l.body is by["mug1"] for l in robot.grabbed[1].noncolliding)      # grabbed-vs-grabbed pairs exist
    res = orc.check_batch(robot, q, want_fk=True)
    env_hit, self_hit = _literal(env, robot, q)
    assert np.array_equal((res["verdict"] & 1) != 0, env_hit)
    assert np.array_equal((res["verdict"] & 2) != 0, self_hit)
    nl = len(robot.links)
    assert res["fk"].shape[1] == nl + 2
    robot.SetDOFValues(q[5].tolist())
    want = [by["mug1"].links[0].t, by["mug2"].links[0].t]
    for k in range(2):
        assert np.allclose(res["fk"][5, nl + k, 0:4], want[k][0], atol=1e-12) and np.allclose(res["fk"][5, nl + k, 4:7], want[k][1], atol=1e-12)
    # the grabbed bodies change verdicts in both directions: new contacts of the mugs, and the env hits of the mugs'
    # old resting places disappear
    assert ((res["verdict"] != 0) & (base == 0)).any()
    fs = res["first_self"]
    assert (fs[(res["verdict"] & 2) != 0].max(axis=1) >= nl).any()               # some first self hit involves a grabbed link
    robot.ReleaseAllGrabbed()
    assert len(robot.grabbed) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("active_only", [False, True])
def test_gpu_equals_oracle_with_grabbed_bodies(active_only):
    from openrave_b200.checker import CO_ActiveDOFs, CudaCollisionChecker
    env, robot, orc, by, hand, fore = _grab_world(lambda e: oracle.OracleChecker(e, "port"))
    env2, robot2, cc, by2, hand2, fore2 = _grab_world(lambda e: CudaCollisionChecker(e, 0))
    for (r, b, h, f) in ((robot, by, hand, fore), (robot2, by2, hand2, fore2)):
        r.Grab(b["mug1"], h)
        r.Grab(b["mug2"], f)
    assert [l.name for l in robot.grabbed[0].noncolliding] == [l.name for l in robot2.grabbed[0].noncolliding]
    assert [l.name for l in robot.grabbed[1].noncolliding] == [l.name for l in robot2.grabbed[1].noncolliding]
    active = None
    if active_only:
        active = robot.GetActiveManipulator().GetArmIndices()
        for r, c in ((robot, orc), (robot2, cc)):
            r.SetActiveDOFs(active)
            c.SetCollisionOptions(CO_ActiveDOFs)
    q = sample_configs(robot, 4000, seed=23, active=active)
    ref = orc.check_batch(robot, q, active=active, want_band=True, want_fk=True)
    got = cc.CheckCollisionBatch(robot2, q, active=active)
    band = (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= 1e-4))
    bad = got["verdict"] != ref["verdict"]
    assert not (bad & ~band).any(), int((bad & ~band).sum())
    ok = ~band & ~bad
    e = ok & ((ref["verdict"] & 1) != 0)
    assert np.array_equal(got["first_env"][e], ref["first_env"][e])
    s = ok & ((ref["verdict"] & 2) != 0)
    assert np.array_equal(got["first_self"][s], ref["first_self"][s])
    nl = len(robot.links)
    assert (ref["first_env"][e][:, 0] >= nl).any() and (ref["first_self"][s].max(axis=1) >= nl).any()
    fk = cc.ForwardKinematicsBatch(robot2, q[:64], active=active)
    assert fk.shape == ref["fk"][:64].shape
    assert np.abs(fk[..., :7] - ref["fk"][:64, :, :7]).max() < 1e-9
    # single-state calls read the bodies' current transforms (grabbed links included)
    for row in q[:12]:
        for r in (robot, robot2):
            r.SetDOFValues(row.tolist(), active)
        assert env.CheckCollision(robot) == env2.CheckCollision(robot2)
        assert robot.CheckSelfCollision() == robot2.CheckSelfCollision()
    robot2.ReleaseAllGrabbed()
    robot.ReleaseAllGrabbed()
    again = cc.CheckCollisionBatch(robot2, q[:500], active=active)
    assert again["query"].info.nlinks == nl                                       # released: the query links are the robot's again
    ref2 = orc.check_batch(robot, q[:500], active=active, want_band=True)
    band2 = (ref2["near"] != 0) | ((ref2["verdict"] != 0) & (ref2["penetration"] <= 1e-4))
    assert not ((again["verdict"] != ref2["verdict"]) & ~band2).any()
