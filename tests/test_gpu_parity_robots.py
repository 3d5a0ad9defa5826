"""CUDA path vs the CPU oracle on the COLLADA robots (BASELINE.json configs[1] PR2/pr2test1, configs[3] HIRO-NX/hironxtable),
through the C ABI (-m gpu).  Same bars as tests/test_gpu_parity.py."""
import numpy as np
import pytest

import oracle
from conftest import ORACLE_KIND, load_scene, sample_configs
from openrave_b200.checker import BATCH_ENV, BATCH_SELF, CO_ActiveDOFs, CudaCollisionChecker

pytestmark = pytest.mark.gpu
BAND = 1e-4
CASES = {"pr2test1": ("rightarm_torso", 2), "hironxtable": (None, 4)}     # (active manipulator, seed) per SURVEY.md §8d


@pytest.fixture(scope="module", params=sorted(CASES))
def world(request):
    scene = request.param
    env = load_scene(scene)
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    orc = oracle.OracleChecker(env, ORACLE_KIND)
    orc.InitEnvironment()
    robot = env.GetRobots()[0]
    manip, seed = CASES[scene]
    active = None
    if manip is not None:
        active = [m for m in robot.GetManipulators() if m.GetName() == manip][0].GetArmIndices()
    return scene, env, robot, cc, orc, active, seed


def classify(ref):
    return (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= BAND))


def test_fk(world):
    scene, env, robot, cc, orc, active, seed = world
    q = sample_configs(robot, 5000, seed=seed, active=active)
    got = cc.ForwardKinematicsBatch(robot, q, active=active)
    ref = orc.check_batch(robot, q, active=active, flags=0, want_fk=True)["fk"]
    dt = np.abs(got[..., 4:7] - ref[..., 4:7]).max()
    dq = np.minimum(np.abs(got[..., 0:4] - ref[..., 0:4]).max(axis=-1), np.abs(got[..., 0:4] + ref[..., 0:4]).max(axis=-1)).max()
    assert dt <= 1e-6 and 2 * dq <= 1e-6, (dt, dq)       # BASELINE.json: 1e-6 m / 1e-6 rad
    assert dt <= 1e-12 and dq <= 1e-12                   # FP64 on the device: same operation order as the reference


def test_verdicts_and_first_pairs(world):
    scene, env, robot, cc, orc, active, seed = world
    n = 3000
    q = sample_configs(robot, n, seed=seed, active=active)
    got = cc.CheckCollisionBatch(robot, q, active=active)
    ref = orc.check_batch(robot, q, active=active, want_band=True)
    band = classify(ref)
    bad = got["verdict"] != ref["verdict"]
    print("%s: configs=%d colliding=%d near-contact=%d verdict-mismatch=%d (outside band: %d) hoisted pairs=%d of %d" % (
        scene, n, int((ref["verdict"] != 0).sum()), int(band.sum()), int(bad.sum()), int((bad & ~band).sum()),
        got["query"].info.npairs_hoisted, got["query"].npairs))
    assert not (bad & ~band).any()
    ok = ~band
    fe = got["first_env"].copy()
    hit = fe[:, 0] >= 0
    fe[hit, 1] = [orc._ids[id(cc.body_from_id(b))] for b in fe[hit, 1]]
    assert (fe[ok] == ref["first_env"][ok]).all()
    assert (got["first_self"][ok] == ref["first_self"][ok]).all()
    assert 0.05 < (ref["verdict"] != 0).mean() < 0.999


def test_activedofs_option(world):
    scene, env, robot, cc, orc, active, seed = world
    if active is None:
        pytest.skip("all DOFs active")
    q = sample_configs(robot, 800, seed=seed + 10, active=active)
    cc.SetCollisionOptions(CO_ActiveDOFs)
    orc.SetCollisionOptions(CO_ActiveDOFs)
    try:
        got = cc.CheckCollisionBatch(robot, q, active=active)
        ref = orc.check_batch(robot, q, active=active, want_band=True)
    finally:
        cc.SetCollisionOptions(0)
        orc.SetCollisionOptions(0)
    bad = got["verdict"] != ref["verdict"]
    assert not (bad & ~classify(ref)).any()


def test_nonadjacent_list_through_gpu(world):
    scene, env, robot, cc, orc, active, seed = world
    assert robot.GetNonAdjacentLinks(0, checker=cc) == robot.GetNonAdjacentLinks(0, checker=orc)


def test_full_size_batch_properties(world):
    """BASELINE.json's full sizes (1M configurations): size-independent properties — a permuted batch gives the permuted
    results, two halves give the whole (chunk boundaries, two-stream pipeline, certificates built mid-way do not matter), a
    random sample agrees with the oracle."""
    scene, env, robot, cc, orc, active, seed = world
    n = 1000000
    q = sample_configs(robot, n, seed=seed, active=active)
    a = cc.CheckCollisionBatch(robot, q, active=active)
    rng = np.random.default_rng(9)
    perm = rng.permutation(n)
    b = cc.CheckCollisionBatch(robot, q[perm], active=active)
    assert np.array_equal(b["verdict"], a["verdict"][perm])
    assert np.array_equal(b["first_env_index"], a["first_env_index"][perm])
    assert np.array_equal(b["first_self_index"], a["first_self_index"][perm])
    h1 = cc.CheckCollisionBatch(robot, q[:n // 2 + 7], active=active)
    h2 = cc.CheckCollisionBatch(robot, q[n // 2 + 7:], active=active)
    assert np.array_equal(np.concatenate([h1["verdict"], h2["verdict"]]), a["verdict"])
    assert np.array_equal(np.concatenate([h1["first_self_index"], h2["first_self_index"]]), a["first_self_index"])
    pick = np.sort(rng.choice(n, 1500, replace=False))
    ref = orc.check_batch(robot, q[pick], active=active, want_band=True)
    band = classify(ref)
    assert not ((a["verdict"][pick] != ref["verdict"]) & ~band).any()
    ok = ~band & (ref["verdict"] == a["verdict"][pick])
    e = ok & ((ref["verdict"] & 1) != 0)
    assert np.array_equal(a["first_env"][pick][e], ref["first_env"][e])
    assert 0.3 < (a["verdict"] != 0).mean() < 0.9
