"""CUDA path vs the CPU oracle on the synthetic clutter workload (BASELINE.json configs[4], openrave_b200/clutter.py), -m gpu.

Small scene: exact verdict / first-pair parity outside the 1e-4 m band.  Full-size scene (200 obstacles x 9 680 triangles,
larger than L2): the oracle still finishes a few hundred configurations in seconds, plus size-independent properties
(determinism, env-only | self-only == env+self, shard == whole)."""
import numpy as np
import pytest

import oracle
from conftest import ORACLE_KIND, sample_configs
from openrave_b200 import clutter
from openrave_b200.checker import BATCH_ENV, BATCH_SELF, CudaCollisionChecker

pytestmark = pytest.mark.gpu
BAND = 1e-4


def make(nob, freq):
    env, robot = clutter.make_clutter_environment(clutter.default_fixture(), nobstacles=nob, freq=freq, seed=5)
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    orc = oracle.OracleChecker(env, ORACLE_KIND)
    orc.InitEnvironment()
    return env, robot, cc, orc, robot.GetActiveDOFIndices()


def compare(robot, cc, orc, active, q):
    got = cc.CheckCollisionBatch(robot, q, active=active)
    ref = orc.check_batch(robot, q, active=active, want_band=True)
    band = (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= BAND))
    bad = got["verdict"] != ref["verdict"]
    print("clutter: configs=%d colliding=%d near-contact=%d verdict-mismatch=%d (outside band: %d)" % (
        len(q), int((ref["verdict"] != 0).sum()), int(band.sum()), int(bad.sum()), int((bad & ~band).sum())))
    assert not (bad & ~band).any()
    ok = ~band
    fe = got["first_env"].copy()
    hit = fe[:, 0] >= 0
    fe[hit, 1] = [orc._ids[id(cc.body_from_id(b))] for b in fe[hit, 1]]
    assert (fe[ok] == ref["first_env"][ok]).all()
    assert (got["first_self"][ok] == ref["first_self"][ok]).all()
    return got, ref


def test_small_clutter_exact():
    env, robot, cc, orc, active = make(24, 6)
    q = sample_configs(robot, 3000, seed=5, active=active)
    got, ref = compare(robot, cc, orc, active, q)
    assert 0.05 < (ref["verdict"] != 0).mean() < 0.95


def test_full_size_clutter():
    env, robot, cc, orc, active = make(200, 22)
    q = sample_configs(robot, 40000, seed=5, active=active)
    compare(robot, cc, orc, active, q[:400])
    bq = cc.bind(robot, active)
    assert bq.info.total_tris > 1900000
    v, fe, fs, _ = bq.check(q, BATCH_ENV | BATCH_SELF)
    v2, fe2, fs2, _ = bq.check(q, BATCH_ENV | BATCH_SELF)
    assert (v == v2).all() and (fe == fe2).all() and (fs == fs2).all()            # deterministic
    ve, fee, _, _ = bq.check(q, BATCH_ENV)
    vs, _, fss, _ = bq.check(q, BATCH_SELF)
    assert ((ve | vs) == v).all() and (fee == fe).all() and (fss == fs).all()      # the two classes are independent
    parts = [bq.check(q[a:b], BATCH_ENV | BATCH_SELF)[0] for a, b in ((0, 13333), (13333, 26667), (26667, 40000))]
    assert (np.concatenate(parts) == v).all()                                      # shards == whole batch
