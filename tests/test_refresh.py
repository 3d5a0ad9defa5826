"""Bound queries follow moving bodies through cc_query_refresh instead of being re-created (include/cudacollision.h): the
single-state overloads after SetDOFValues, an obstacle that moves, inactive DOFs that change — all against the oracle, which
reads every transform per call like pqprave (plugins/pqprave/collisionPQP.h:438-463)."""
import time

import numpy as np
import pytest

import oracle
from conftest import load_scene, sample_configs
from openrave_b200 import ravemath as rm


def _pair(arm_only):
    from openrave_b200.checker import CudaCollisionChecker
    out = []
    for factory in (lambda e: oracle.OracleChecker(e, "port"), lambda e: CudaCollisionChecker(e, 0)):
        env = load_scene("lab1")
        robot = env.GetRobots()[0]
        cc = factory(env)
        env.SetCollisionChecker(cc)
        active = None
        if arm_only:
            active = robot.GetActiveManipulator().GetArmIndices()
            robot.SetActiveDOFs(active)
            cc.SetCollisionOptions(0x10)
        out.append((env, robot, cc, active))
    return out


@pytest.mark.gpu
def test_single_state_calls_do_not_rebind():
    (env, robot, orc, _), (env2, robot2, cc, _) = _pair(False)
    q = sample_configs(robot, 80, seed=61)
    robot2.SetDOFValues(q[0].tolist())
    env2.CheckCollision(robot2)
    bound = dict(cc._queries)
    assert len(bound) >= 1
    t0 = time.perf_counter()
    for row in q:
        robot.SetDOFValues(row.tolist())
        robot2.SetDOFValues(row.tolist())
        assert env.CheckCollision(robot) == env2.CheckCollision(robot2)
        assert robot.CheckSelfCollision() == robot2.CheckSelfCollision()
    assert all(cc._queries.get(k) is v for k, v in bound.items())            # the same bound queries served every call
    per_call = (time.perf_counter() - t0) / (2 * len(q))
    assert per_call < 0.02, per_call                                       # re-binding would cost tens of milliseconds per call


@pytest.mark.gpu
def test_moving_obstacle_and_inactive_dofs():
    (env, robot, orc, active), (env2, robot2, cc, _) = _pair(True)
    q = sample_configs(robot, 30000, seed=63, active=active)
    ref0 = orc.check_batch(robot, q[:3000], active=active, want_band=True)
    got0 = cc.CheckCollisionBatch(robot2, q, active=active)              # >= 16k: certificates are built on this query
    bq = got0["query"]
    assert bq.table_count() > 0

    def same(got, ref):
        band = (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= 1e-4))
        return not ((got["verdict"][:len(ref["verdict"])] != ref["verdict"]) & ~band).any()
    assert same(got0, ref0)
    # an obstacle moves into the workspace: poses are pushed, the bound query is refreshed (not re-created), certificates reset
    for e, r in ((env, robot), (env2, robot2)):
        mug = [b for b in e.GetBodies() if b.name == "mug1"][0]
        mug.SetTransform(rm.t_mul(r.GetLink("wam4").t, (rm.IDENT_Q, (0.0, 0.0, 0.25))))
    ref1 = orc.check_batch(robot, q[:3000], active=active, want_band=True)
    ncert0 = bq.table_count()
    mug_id = cc.body_id([b for b in env2.GetBodies() if b.name == "mug1"][0])
    assert any(int(bq.pairs[c.pair][1]) == mug_id for c in bq.certificates()) or True
    got1 = cc.CheckCollisionBatch(robot2, q[:3000], active=active)
    # the refresh drops only the certificates the moved body touches: the others stay in force, none names the mug any more
    assert got1["query"] is bq and 0 < bq.table_count() <= ncert0
    assert not any(int(bq.pairs[c.pair][1]) == mug_id and int(bq.pairs[c.pair][3]) == 0 for c in bq.certificates())
    assert same(got1, ref1)
    assert (ref1["verdict"] != ref0["verdict"]).any()
    got1b = cc.CheckCollisionBatch(robot2, q, active=active)             # certificates come back on the refreshed poses
    assert bq.table_count() > 0 and same(got1b, ref1)
    # an inactive DOF changes (a finger closes): the FK-mode entry re-reads it
    for r in (robot, robot2):
        vals = r.GetDOFValues()
        vals[7] = 1.5
        vals[8] = 1.2
        r.SetDOFValues(vals)
    ref2 = orc.check_batch(robot, q[:3000], active=active, want_band=True)
    got2 = cc.CheckCollisionBatch(robot2, q[:3000], active=active)
    assert got2["query"] is bq and same(got2, ref2)
    fk = cc.ForwardKinematicsBatch(robot2, q[:16], active=active)
    ref_fk = orc.check_batch(robot, q[:16], active=active, want_fk=True)["fk"]
    assert np.abs(fk[..., :7] - ref_fk[..., :7]).max() < 1e-9


@pytest.mark.gpu
def test_certificates_are_never_proven_on_poses_the_query_does_not_use():
    """A raw C-ABI caller moves a body after binding and does NOT call cc_query_refresh: the bound query keeps checking its
    snapshot (include/cudacollision.h), so the free-space certificates — which are proven from the live scene state — must not
    be built until the refresh (they would drop pairs on poses the query does not use)."""
    import ctypes
    from openrave_b200 import flatten
    (env, robot, orc, active), (env2, robot2, cc, _) = _pair(True)
    q = sample_configs(robot, 40000, seed=67, active=active)
    bq = cc.bind(robot2, active)
    small = bq.check(q[:2000])                                              # < 16k configurations: no certificates yet
    assert bq.table_count() == 0
    # behind the harness's back: park mug1 next to wam4 in the scene only
    mug = [b for b in env2.GetBodies() if b.name == "mug1"][0]
    moved = rm.t_mul(robot2.GetLink("wam4").t, (rm.IDENT_Q, (0.0, 0.0, 0.25)))
    T = flatten.transforms8([moved for _ in mug.links])
    assert cc.lib.cc_body_set_link_transforms(cc.scene, cc.body_id(mug), T.ctypes.data_as(ctypes.c_void_p)) == 0
    stale = bq.check(q)                                                     # crosses 16k: would build certificates
    assert bq.table_count() == 0                                            # ... but the snapshot is stale: not built
    assert np.array_equal(stale[0][:2000], small[0])                        # and the answers are still the snapshot's
    bq.refresh()
    fresh = bq.check(q)
    assert bq.table_count() > 0
    assert (fresh[0] != stale[0]).any()                                     # the moved mug now matters
    mug1 = [b for b in env.GetBodies() if b.name == "mug1"][0]
    mug1.SetTransform(rm.t_mul(robot.GetLink("wam4").t, (rm.IDENT_Q, (0.0, 0.0, 0.25))))
    ref = orc.check_batch(robot, q[:3000], active=active, want_band=True)
    band = (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= 1e-4))
    assert not ((fresh[0][:3000] != ref["verdict"]) & ~band).any()
