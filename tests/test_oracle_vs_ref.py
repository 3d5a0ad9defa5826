"""Pin the restated PQP arithmetic (oracle/liboracle.so) against the UNMODIFIED vendored PQP (oracle/_ref).

Same link meshes, same poses -> identical contact id lists and identical NumBVTests/NumTriTests
(pqp/PQP.cpp:489,500), which also pins the BVH build (Build.cpp, BV.cpp, Meigen).  not gpu.
"""
import numpy as np
import pytest

import oracle
from conftest import have_ref_oracle, load_scene, sample_configs

pytestmark = pytest.mark.skipif(not have_ref_oracle(), reason="oracle/_ref not built (needs /root/reference at build time)")


def test_pair_queries_identical():
    env = load_scene("lab1")
    port = oracle.OracleChecker(env, "port")
    ref = oracle.OracleChecker(env, "reference")
    port.InitEnvironment()
    ref.InitEnvironment()
    robot = env.GetRobots()[0]
    env.SetCollisionChecker(port)
    q = list(sample_configs(robot, 6, seed=7))
    q.append(np.array([-8.44575603e-02, 1.48528347e+00, -5.09108824e-08, 6.48108822e-01, -4.57571203e-09, -1.04008750e-08,
                       7.26855048e-10, 5.50807826e-08, 5.50807826e-08, -1.90689327e-08, 0.0]))   # test_collision.py:187
    others = [l for b in env.GetBodies() if b is not robot for l in b.links]
    ncontacts = 0
    for row in q:
        robot.SetDOFValues(row)
        for rl in robot.links:
            for ol in others:
                n1, c1, k1 = port.collide_pair(rl, ol)
                n2, c2, k2 = ref.collide_pair(rl, ol)
                assert n1 == n2
                assert (c1 == c2).all(), (rl.name, ol.body.name, c1, c2)
                assert (k1 == k2).all()
                ncontacts += n1
    assert ncontacts > 0


def test_batch_identical():
    env = load_scene("lab1")
    port = oracle.OracleChecker(env, "port")
    ref = oracle.OracleChecker(env, "reference")
    port.InitEnvironment()
    ref.InitEnvironment()
    robot = env.GetRobots()[0]
    q = sample_configs(robot, 150, seed=11)
    a = port.check_batch(robot, q, report=True, want_fk=True)
    b = ref.check_batch(robot, q, report=True, want_fk=True)
    assert (a["verdict"] == b["verdict"]).all()
    assert (a["first_env"] == b["first_env"]).all()
    assert (a["first_self"] == b["first_self"]).all()
    assert (a["counters"] == b["counters"]).all()
    assert (a["env_pairs"] == b["env_pairs"]).all()
    assert np.array_equal(a["fk"], b["fk"])
    assert 0 < (a["verdict"] != 0).sum() < len(q)


@pytest.mark.parametrize("scene,manip,n,seed", [("pr2test1", "rightarm_torso", 400, 2), ("hironxtable", None, 400, 4),
                                                ("wamtest1", "arm", 300, 3)])
def test_batch_identical_on_the_other_robots(scene, manip, n, seed):
    """PR2 (88 links, 901 self pairs) and HIRO-NX: the restatement and the unmodified PQP agree on verdicts, first pairs, every
    reported env pair AND the per-configuration NumBVTests / NumTriTests — which also pins the restated model build on the
    COLLADA meshes.  The GPU parity tests may therefore use either oracle build."""
    env = load_scene(scene)
    robot = env.GetRobots()[0]
    active = None
    if manip is not None:
        active = [m for m in robot.GetManipulators() if m.GetName() == manip][0].GetArmIndices()
        robot.SetActiveDOFs(active)
    port = oracle.OracleChecker(env, "port")
    ref = oracle.OracleChecker(env, "reference")
    port.InitEnvironment()
    ref.InitEnvironment()
    q = sample_configs(robot, n, seed=seed, active=active)
    a = port.check_batch(robot, q, active=active, report=True, want_fk=True, want_band=True)
    b = ref.check_batch(robot, q, active=active, report=True, want_fk=True, want_band=True)
    for k in ("verdict", "first_env", "first_self", "counters", "env_pairs", "near"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["fk"], b["fk"])
    assert np.allclose(a["penetration"], b["penetration"], rtol=0, atol=1e-15)
    assert 0 < (a["verdict"] != 0).sum() < len(q)


def test_parity_record_of_identical_answers_is_clean():
    """oracle/parity.py (what the full-size sweeps and bench.py report): the restated oracle plays the CUDA path's role against
    the reference-linked oracle on all cores; identical answers give a clean record, an injected verdict flip outside the band
    is counted as such."""
    from oracle import parity
    import bench
    env, robot, _, seed = bench.load_world("wam_lab1_100k")
    q = bench.sample(robot, 100000, seed)[:600]
    ref = parity.oracle_pass("wam_lab1_100k", q, kind="reference", workers=2)
    port = oracle.OracleChecker(env, "port")
    port.InitEnvironment()
    mine = port.check_batch(robot, q)
    names = {bid: b.name for bid, b in port._bodies_by_id().items()}
    rec = parity.compare(mine, ref, lambda b: names[b])
    assert rec["compared"] == 600 and rec["identical"] == 600 and rec["mismatch_outside_band"] == 0
    assert 0 < rec["colliding"] < 600
    band = parity.band_mask(ref)
    k = int(np.nonzero(~band)[0][0])
    mine["verdict"] = mine["verdict"].copy()
    mine["verdict"][k] ^= 1
    rec = parity.compare(mine, ref, lambda b: names[b])
    assert rec["mismatch_outside_band"] == 1 and rec["outside_band_indices"] == [k]
