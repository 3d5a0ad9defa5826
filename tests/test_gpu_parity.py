"""CUDA path vs the CPU oracle on the same seeded inputs, through the C ABI (-m gpu).

Bars (BASELINE.json north_star): link transforms within 1e-6 m / 1e-6 rad of KinBody::SetDOFValues; verdicts and
reported link pairs identical on every configuration outside the 1e-4 m contact band; near-contact
configurations are counted and reported.
"""
import numpy as np
import pytest

import oracle
from conftest import ORACLE_KIND, load_scene, sample_configs
from openrave_b200.checker import (BATCH_ALL_PAIRS, BATCH_ENV, BATCH_SELF, CO_ActiveDOFs, CudaCollisionChecker)
from openrave_b200.environment import CollisionReport

pytestmark = pytest.mark.gpu

BAND = 1e-4
KNOWN_COLLISION = [-8.44575603e-02, 1.48528347e+00, -5.09108824e-08, 6.48108822e-01, -4.57571203e-09, -1.04008750e-08,
                   7.26855048e-10, 5.50807826e-08, 5.50807826e-08, -1.90689327e-08, 0.00000000e+00]


@pytest.fixture(scope="module")
def world():
    env = load_scene("lab1")
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    orc = oracle.OracleChecker(env, ORACLE_KIND)          # unmodified PQP when oracle/_ref travelled with the snapshot
    orc.InitEnvironment()
    return env, env.GetRobots()[0], cc, orc


def quat_angle(qa, qb):
    d = np.abs(np.sum(qa * qb, axis=-1)).clip(0, 1)
    return 2 * np.arccos(d)


def test_fk_matches_setdofvalues(world):
    env, robot, cc, orc = world
    q = sample_configs(robot, 20000, seed=1)
    got = cc.ForwardKinematicsBatch(robot, q)
    ref = orc.check_batch(robot, q[:20000], flags=0, want_fk=True)["fk"]
    dt = np.abs(got[..., 4:7] - ref[..., 4:7]).max()
    da = quat_angle(got[..., 0:4], ref[..., 0:4]).max()
    assert dt <= 1e-6 and da <= 1e-6, (dt, da)      # tolerance stated by BASELINE.json
    assert dt <= 1e-12 and da <= 1e-7               # FP64 on the device: far inside it (acos limits the angle check)
    # values beyond the limits are clamped like SetDOFValues does (kinbody.cpp:2149-2224)
    lo, hi = robot.GetDOFLimits()
    qq = np.array([np.array(hi) + 0.3, np.array(lo) - 0.3])
    g2 = cc.ForwardKinematicsBatch(robot, qq)
    r2 = orc.check_batch(robot, qq, flags=0, want_fk=True)["fk"]
    assert np.abs(g2 - r2).max() <= 1e-12


def classify(ref):
    """near-contact: free but closer than BAND, or colliding with penetration measure <= BAND."""
    return (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= BAND))


def test_verdicts_and_first_pairs_match_oracle(world):
    env, robot, cc, orc = world
    n = 4000
    q = sample_configs(robot, n, seed=1)            # C1's seed; a prefix of the 100k batch
    got = cc.CheckCollisionBatch(robot, q)
    ref = orc.check_batch(robot, q, want_band=True)
    band = classify(ref)
    bad = got["verdict"] != ref["verdict"]
    print("configs=%d colliding=%d near-contact=%d verdict-mismatch=%d (outside band: %d)" % (
        n, int((ref["verdict"] != 0).sum()), int(band.sum()), int(bad.sum()), int((bad & ~band).sum())))
    assert not (bad & ~band).any()
    ok = ~band
    # reported pairs: the first colliding pair in pqprave's iteration order
    fe = got["first_env"].copy()
    hit = fe[:, 0] >= 0
    fe[hit, 1] = [orc._ids[id(cc.body_from_id(b))] for b in fe[hit, 1]]
    assert (fe[ok] == ref["first_env"][ok]).all()
    assert (got["first_self"][ok] == ref["first_self"][ok]).all()
    assert 0.05 < (ref["verdict"] != 0).mean() < 0.999


def test_all_pairs_mask_matches_report_mode(world):
    """CO_AllLinkCollisions: the complete pair mask against the reference's report-mode pair list, pair by pair.  A pair may
    differ only if THAT pair sits inside the contact band (oracle: free within 1e-4 m, or colliding with a penetration measure
    <= 1e-4 m); outside the band the tolerance is zero."""
    env, robot, cc, orc = world
    n = 3000
    q = sample_configs(robot, n, seed=21)
    got = cc.CheckCollisionBatch(robot, q, flags=BATCH_ENV, all_pairs=True)
    ref = orc.check_batch(robot, q, flags=BATCH_ENV, report=True)
    bq = got["query"]
    npe = bq.info.npairs_env
    # decode both sides into [n, npairs_env] bit matrices over the product's pair indices
    key = {(int(l1), orc._ids[id(cc.body_from_id(int(ob)))], int(l2)): p for p, (l1, ob, l2, _) in enumerate(bq.pairs[:npe])}
    mine = ((got["mask"][:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(n, -1)[:, :npe].astype(bool)
    theirs = np.zeros((n, npe), dtype=bool)
    for c, l1, ob, l2 in ref["env_pairs"]:
        theirs[c, key[(int(l1), int(ob), int(l2))]] = True      # a KeyError here = a pair the product does not even enumerate
    assert theirs.sum() > 1000
    diff = np.argwhere(mine != theirs)
    nband = 0
    for c, p in diff:
        l1, ob, l2, _ = bq.pairs[p]
        cls = orc.pair_band(robot, q[c:c + 1], [(int(l1), orc._ids[id(cc.body_from_id(int(ob)))], int(l2))])[0, 0]
        assert cls in (1, 2), "pair %d of configuration %d differs outside the contact band (class %d)" % (p, c, cls)
        nband += 1
    print("all-pairs mask: %d configurations x %d env pairs, %d colliding pairs, %d differing (all inside the band)" % (
        n, npe, int(theirs.sum()), nband))
    assert nband <= n // 100          # the band is 1e-4 m wide: more than a handful means something else is wrong
    # the first pair of the complete mask is the first pair of the early-exit flavour
    first = cc.CheckCollisionBatch(robot, q, flags=BATCH_ENV)
    has = mine.any(axis=1)
    assert np.array_equal(has, (first["verdict"] & 1).astype(bool))
    assert np.array_equal(mine.argmax(axis=1)[has], first["first_env_index"][has])


def test_hoisted_pairs_and_activedofs(world):
    env, robot, cc, orc = world
    arm = robot.GetActiveManipulator().GetArmIndices()
    q = sample_configs(robot, 1000, seed=5, active=arm)
    for opts in (0, CO_ActiveDOFs):
        cc.SetCollisionOptions(opts)
        orc.SetCollisionOptions(opts)
        got = cc.CheckCollisionBatch(robot, q, active=arm)
        ref = orc.check_batch(robot, q, active=arm, want_band=True)
        band = classify(ref)
        bad = got["verdict"] != ref["verdict"]
        assert not (bad & ~band).any(), opts
        info = got["query"].info
        if opts == 0:
            assert info.npairs_hoisted > 0       # segway / wam0 vs the static environment cannot change with the arm
        else:
            assert info.npairs_hoisted == 0      # CO_ActiveDOFs drops the links no active DOF moves (collisionPQP.h:710-735)
    cc.SetCollisionOptions(0)
    orc.SetCollisionOptions(0)


def test_single_state_overloads(world):
    env, robot, cc, orc = world
    robot.SetDOFValues([0.0] * 11)
    assert not env.CheckCollision(robot)
    assert env.CheckCollision(env.GetKinBody('mug1'))
    assert not robot.CheckSelfCollision()
    robot.SetDOFValues(KNOWN_COLLISION)
    rep = CollisionReport()
    assert env.CheckCollision(robot, report=rep)
    rep2 = CollisionReport()
    assert orc.CheckCollision(robot, report=rep2)
    assert [(a.name, b.body.name, b.name) for a, b in rep.vLinkColliding] == \
           [(a.name, b.body.name, b.name) for a, b in rep2.vLinkColliding]
    # link/link and link/body overloads (test_collision.py:85-91 style)
    f02 = robot.GetLink('Finger0-2')
    seg = env.GetKinBody('segway')
    assert env.CheckCollision(f02, seg)
    assert env.CheckCollision(f02, seg.links[0])
    assert not env.CheckCollision(robot.GetLink('wam1'), seg)
    robot.SetDOFValues([0.0] * 11)
    # non-adjacent list computed through the GPU (kinbody.cpp:5563-5577) equals the oracle's
    assert robot.GetNonAdjacentLinks(0, checker=cc) == robot.GetNonAdjacentLinks(0, checker=orc)


def test_pose_mode_equals_fk_mode(world):
    env, robot, cc, orc = world
    q = sample_configs(robot, 600, seed=9)
    bq = cc.bind(robot)
    T = bq.fk(q)
    v1, fe1, fs1, _ = bq.check(q)
    v2, fe2, fs2, _ = bq.check_poses(T)
    assert (v1 == v2).all() and (fe1 == fe2).all() and (fs1 == fs2).all()


def test_empty_and_tiny_batches(world):
    env, robot, cc, orc = world
    bq = cc.bind(robot)
    v, fe, fs, _ = bq.check(np.zeros((0, 11)))
    assert v.shape == (0,)
    v, fe, fs, _ = bq.check(np.zeros((1, 11)))
    assert v[0] == 0 and fe[0] == -1 and fs[0] == -1
    v, _, _, _ = bq.check(np.array([KNOWN_COLLISION]))
    assert v[0] & 1


def test_determinism(world):
    env, robot, cc, orc = world
    q = sample_configs(robot, 3000, seed=13)
    a = cc.CheckCollisionBatch(robot, q)
    b = cc.CheckCollisionBatch(robot, q)
    assert (a["verdict"] == b["verdict"]).all()
    assert (a["first_env_index"] == b["first_env_index"]).all()
    assert (a["first_self_index"] == b["first_self_index"]).all()


@pytest.mark.gpu
def test_first_hit_mode_keeps_the_planner_return_code(world):
    """CC_BATCH_FIRST_HIT = the report == NULL path: the env bit is exact, the self bit is exact whenever the env bit is clear,
    i.e. the code DynamicsCollisionConstraint::_CheckState returns (env first, then self, planningutils.cpp:2523-2548)."""
    from openrave_b200.checker import BATCH_ENV, BATCH_SELF, BATCH_FIRST_HIT
    env, robot, cc, orc = world[:4]
    q = sample_configs(robot, 30000, seed=71)
    full = cc.CheckCollisionBatch(robot, q)["verdict"]
    fast = cc.CheckCollisionBatch(robot, q, flags=BATCH_ENV | BATCH_SELF | BATCH_FIRST_HIT)["verdict"]
    assert np.array_equal(fast & 1, full & 1)
    free_env = (full & 1) == 0
    assert np.array_equal(fast[free_env], full[free_env])
    assert ((fast & 2) <= (full & 2)).all()                 # a reported self contact is a real one
    code = lambda v: np.where(v & 1, 1, np.where(v & 2, 2, 0))
    assert np.array_equal(code(fast), code(full))
    for n in (1, 50, 700):                                  # split kernel
        f2 = cc.CheckCollisionBatch(robot, q[:n], flags=BATCH_ENV | BATCH_SELF | BATCH_FIRST_HIT)["verdict"]
        assert np.array_equal(code(f2), code(full[:n]))
