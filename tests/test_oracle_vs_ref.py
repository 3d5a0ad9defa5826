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
