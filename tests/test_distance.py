"""CO_Distance: minimum separation per configuration.

pqprave answers distance queries with PQP_Distance at rel_err = 200 / abs_err = 0.001 (plugins/pqprave/collisionPQP.h:48-49,
622-648), an approximate search; the product computes the exact minimum (both tolerances 0).  Pins:
  CPU  the oracle's restated branch and bound == unmodified PQP_Distance(…, 0, 0) to 1e-12, and the exact values sit at or
       below the reference's own golden numbers for lab1 (test/test_collision.py:223-251) with the same nearest link pairs;
  GPU  the CUDA kernel == the oracle to 2e-5 m (FP32 evaluation), same pair, closest points |p1 - p2| == distance,
       distance 0 <=> collision verdict, the search radius clamps, the single-state report path."""
import numpy as np
import pytest

import oracle
from conftest import have_ref_oracle, load_scene, sample_configs

TOL = 2e-5


def _lab1(kind_or_factory, arm_only=False):
    env = load_scene("lab1")
    robot = env.GetRobots()[0]
    cc = oracle.OracleChecker(env, kind_or_factory) if isinstance(kind_or_factory, str) else kind_or_factory(env)
    env.SetCollisionChecker(cc)
    active = None
    if arm_only:
        active = robot.GetActiveManipulator().GetArmIndices()
        robot.SetActiveDOFs(active)
        cc.SetCollisionOptions(0x10)           # CO_ActiveDOFs: the segway base (2.7 mm above the floor) drops out
    return env, robot, cc, active


@pytest.mark.skipif(not have_ref_oracle(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_restated_distance_equals_reference_pqp():
    for arm_only in (False, True):
        _, robot, port, active = _lab1("port", arm_only)
        _, robot2, ref, _ = _lab1("reference", arm_only)
        q = sample_configs(robot, 40, seed=31, active=active)
        a = port.distance_batch(robot, q, active=active)
        b = ref.distance_batch(robot2, q, active=active)
        assert np.abs(a["distance"] - b["distance"]).max() < 1e-12
        assert np.array_equal(a["pair"], b["pair"])
        v = port.check_batch(robot, q, active=active)["verdict"]
        assert np.array_equal(a["distance"] == 0, v != 0)


def test_exact_distance_against_the_reference_goldens():
    """test/test_collision.py:238-250: pqp reports 0.0027169 (segway, floorwalls) and, with CO_ActiveDOFs on the arm,
    0.29193971893003506 (wam1, pole) for lab1 at its load pose; the exact minimum can only be smaller or equal."""
    env, robot, orc, _ = _lab1("port")
    by_id = orc._bodies_by_id()
    from openrave_b200 import flatten
    T = flatten.transforms8(robot.GetLinkTransformations())[None]
    r = orc.distance_batch(robot, None, flags=1, poses=T)
    l, ob, ol = r["pair"][0]
    assert (robot.links[l].name, by_id[ob].name) == ("segway", "floorwalls")
    assert abs(r["distance"][0] - 0.0027169) < 1e-4
    env, robot, orc, active = _lab1("port", arm_only=True)
    by_id = orc._bodies_by_id()
    r = orc.distance_batch(robot, None, active=active, flags=1, poses=T)
    l, ob, ol = r["pair"][0]
    assert (robot.links[l].name, by_id[ob].name) == ("wam1", "pole")
    assert 0.28 < r["distance"][0] <= 0.29193971893003506 + 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("arm_only", [False, True])
def test_gpu_distance_equals_oracle(arm_only):
    from openrave_b200.checker import CudaCollisionChecker
    _, robot, orc, active = _lab1("port", arm_only)
    _, robot2, cc, _ = _lab1(lambda e: CudaCollisionChecker(e, 0), arm_only)
    q = sample_configs(robot, 3000, seed=37, active=active)
    ref = orc.distance_batch(robot, q, active=active)
    got = cc.DistanceBatch(robot2, q, active=active)
    err = np.abs(got["distance"] - ref["distance"])
    assert err.max() < TOL, err.max()
    # the same pair attains it, except where two pairs tie within the FP32 error
    sep = ref["distance"] > 1e-4               # colliding configurations tie at 0 over every colliding pair
    same = (got["pair"] == ref["pair"]).all(axis=1)
    assert same[sep].mean() > 0.98, same[sep].mean()
    seg = np.linalg.norm(got["points"][:, 0:3] - got["points"][:, 3:6], axis=1)
    assert np.abs(seg[sep] - got["distance"][sep]).max() < TOL
    assert len(np.unique(np.round(ref["distance"], 6))) > (100 if arm_only else 3)
    # distance 0 <=> the collision entry reports a contact (outside the contact band)
    v = cc.CheckCollisionBatch(robot2, q, active=active)["verdict"]
    clear = ref["distance"] > 1e-4
    assert not v[clear].any() and (v[ref["distance"] == 0] != 0).all()
    # search radius
    capped = cc.DistanceBatch(robot2, q[:500], active=active, max_distance=0.05)
    assert np.abs(capped["distance"] - np.minimum(ref["distance"][:500], 0.05)).max() < TOL
    # PQP_Distance's tolerances: exact <= r <= max(exact + abs_err, exact * (1 + rel_err)); pqprave's own setting last
    for rel, ab in ((0.5, 0.0), (0.0, 0.005), (200.0, 0.001)):
        ap = cc.DistanceBatch(robot2, q[:1000], active=active, rel_err=rel, abs_err=ab)["distance"]
        ex = ref["distance"][:1000]
        assert (ap >= ex - TOL).all() and (ap <= np.maximum(ex + ab, ex * (1 + rel)) + TOL).all()
    # env only / self only
    for flags in (1, 2):
        r1 = orc.distance_batch(robot, q[:300], active=active, flags=flags)
        g1 = cc.DistanceBatch(robot2, q[:300], active=active, flags=flags)
        assert np.abs(g1["distance"] - r1["distance"]).max() < TOL


@pytest.mark.gpu
def test_gpu_distance_pr2_and_report_path():
    from openrave_b200.checker import CO_ActiveDOFs, CO_Distance, CudaCollisionChecker
    from openrave_b200.environment import CollisionReport
    env = load_scene("pr2test1")
    robot = env.GetRobots()[0]
    arm = [m for m in robot.GetManipulators() if m.GetName() == "rightarm_torso"][0].GetArmIndices()
    robot.SetActiveDOFs(arm)
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    orc.SetCollisionOptions(CO_ActiveDOFs)
    env2 = load_scene("pr2test1")
    robot2 = env2.GetRobots()[0]
    robot2.SetActiveDOFs(arm)
    cc = CudaCollisionChecker(env2, 0)
    env2.SetCollisionChecker(cc)
    assert cc.SetCollisionOptions(CO_ActiveDOFs | CO_Distance)
    q = sample_configs(robot, 600, seed=41, active=arm)
    ref = orc.distance_batch(robot, q, active=arm)
    got = cc.DistanceBatch(robot2, q, active=arm)
    assert np.abs(got["distance"] - ref["distance"]).max() < TOL
    # CheckCollision(robot, report) with CO_Distance: minDistance + the closest pair + one contact (collisionPQP.h:622-648)
    from openrave_b200 import flatten
    for row in q[:8]:
        robot2.SetDOFValues(row.tolist(), arm)
        robot.SetDOFValues(row.tolist(), arm)
        rep = CollisionReport()
        hit = env2.CheckCollision(robot2, report=rep)
        want = orc.distance_batch(robot, None, active=arm, flags=1, poses=flatten.transforms8(robot.GetLinkTransformations())[None])
        if hit:
            assert rep.minDistance == 0
        else:
            assert abs(rep.minDistance - want["distance"][0]) < TOL
            pos, norm, depth = rep.contacts[0]
            assert abs(-depth - rep.minDistance) < TOL and abs(np.linalg.norm(norm) - 1) < 1e-6
            assert rep.plink1.name == robot.links[want["pair"][0][0]].name


@pytest.mark.skipif(not have_ref_oracle(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_tolerance_restatement_equals_reference_pqp():
    """PQP_Tolerance of the unmodified library == (exact distance <= tolerance) of the restatement, pair by pair."""
    _, robot, port, active = _lab1("port", True)
    _, robot2, ref, _ = _lab1("reference", True)
    by = port._bodies_by_id()
    triples = [(l.index, bid, ol.index) for l in robot.links[1:9] for bid, b in by.items() if b is not robot for ol in b.links][:60]
    q = sample_configs(robot, 25, seed=43, active=active)
    tol = 0.05
    d, _ = port.pair_queries(robot, q, triples, active=active, cap=1.0, tolerance=tol)
    _, w = ref.pair_queries(robot2, q, triples, active=active, cap=0.0, tolerance=tol)
    clear = np.abs(d - tol) > 1e-9
    assert np.array_equal((d <= tol)[clear], w[clear]) and w.any() and not w.all()


@pytest.mark.gpu
def test_gpu_tolerance_mask_equals_oracle():
    from openrave_b200.checker import CO_ActiveDOFs, CO_UseTolerance, CudaCollisionChecker
    from openrave_b200.environment import CollisionReport
    _, robot, orc, active = _lab1("port", True)
    env2, robot2, cc, _ = _lab1(lambda e: CudaCollisionChecker(e, 0), True)
    q = sample_configs(robot, 300, seed=47, active=active)
    tol = 0.03
    bq = cc.bind(robot2, active)
    w = bq.within_tolerance(tol, q)
    # the query's pairs as (query link, other body, other link) triples on the oracle's body ids (same order of bodies)
    triples = [(int(a), int(ob), int(j)) for a, ob, j, _ in bq.pairs]
    d, _ = orc.pair_queries(robot, q, triples, active=active, cap=tol * 2, tolerance=tol)
    clear = np.abs(d - tol) > TOL
    assert np.array_equal(w[clear], (d <= tol)[clear])
    assert w.any(axis=1).mean() > 0.2 and not w.all()
    # the report path: numWithinTol = other bodies with a pair within tolerance (collisionPQP.h:443-497)
    cc.SetCollisionOptions(CO_ActiveDOFs)
    cc.SetTolerance(tol)
    assert cc.GetCollisionOptions() & CO_UseTolerance
    bq = cc.bind(robot2, active)
    for row, wrow in zip(q[:6], (d <= tol)[:6]):
        robot2.SetDOFValues(row.tolist(), active)
        rep = CollisionReport()
        env2.CheckCollision(robot2, report=rep)
        want = len(set(t[1] for t, inside, cls in zip(triples, wrow, bq.pairs[:, 3]) if inside and cls == 0))
        assert rep.numWithinTol == want
