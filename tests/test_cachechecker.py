"""`CacheChecker` semantics on the batched entry (openrave_b200/cachechecker.py <- plugins/configurationcache/cachechecker.cpp,
configurationcachetree.cpp): thresholds, insertion separation, hits / misses, and the one-inner-batch miss path.  CPU, with the
oracle as the wrapped checker (the reference's own test, test/test_configurationcache.py:283-292, asserts a hit rate > 0.9 on
repeated planning queries — the same property is checked here on repeated batches)."""
import numpy as np

import oracle
from conftest import load_scene, sample_configs
from openrave_b200.cachechecker import CacheCollisionChecker, ConfigurationCache
from test_edge_check import OracleBatchAdapter


class _Inner(OracleBatchAdapter):
    def __init__(self, orc):
        OracleBatchAdapter.__init__(self, orc)
        self.calls = 0
        self.configs = 0

    def CheckCollisionBatch(self, robot, q, flags=3, active=None, all_pairs=False):
        self.calls += 1
        self.configs += len(q)
        return OracleBatchAdapter.CheckCollisionBatch(self, robot, q, flags=flags, active=active)

    def __getattr__(self, name):
        return getattr(self.orc, name)


def _world():
    env = load_scene("wamtest1")
    robot = env.GetRobots()[0]
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    return env, robot, orc


def test_cache_rules():
    env, robot, orc = _world()
    c = ConfigurationCache(robot)
    res = np.array([robot.joints[d].resolution for d in c.active])
    assert np.allclose(c.weights, 1.0 / res)                                  # configurationcachetree.cpp:1194-1201
    q0 = np.zeros(len(c.active))
    assert c.CheckCollisionMany(q0[None])[0][0] == -1                         # empty cache: miss
    assert c.InsertConfiguration(q0, None)
    step = res / np.sqrt(len(res))                                            # one weighted unit of distance along the diagonal
    assert c.CheckCollisionMany((q0 + 0.19 * step)[None])[0][0] == 0          # inside _freespacethresh = 0.2
    assert c.CheckCollisionMany((q0 + 0.21 * step)[None])[0][0] == -1
    assert not c.InsertConfiguration(q0 + 0.09 * step, None)                  # within 0.2 * 0.5 of a node: refused
    assert c.InsertConfiguration(q0 + 3.0 * step, ("a", "b"))
    ret, node = c.CheckCollisionMany(np.array([q0 + 2.1 * step, q0 + 4.1 * step]))
    assert ret[0] == 1 and c.pairs[node[0]] == ("a", "b") and ret[1] == -1    # inside / outside _collisionthresh = 1.0
    assert c.CheckCollisionMany((q0 + 0.15 * step)[None])[0][0] == 0          # free node closer than its threshold, collision node far
    assert not c.InsertConfiguration(q0 + 3.4 * step, ("c", "d"))             # within 1.0 * 0.5 of the collision node
    assert c.GetNumKnownNodes() == 2


def test_batched_misses_go_to_the_inner_checker_once_and_repeats_hit():
    env, robot, orc = _world()
    inner = _Inner(orc)
    # tight thresholds: a cache answer is then (almost always) the exact answer, so the comparison below is meaningful
    cc = CacheCollisionChecker(inner, robot, collisionthresh=0.05, freespacethresh=0.05)
    act = robot.GetActiveDOFIndices()
    q = sample_configs(robot, 300, seed=4, active=act)
    exact = orc.check_batch(robot, q, active=act)["verdict"]
    r1 = cc.CheckCollisionBatch(robot, q)
    assert inner.calls == 1 and inner.configs == 300                           # empty caches: everything misses, ONE inner batch
    assert np.array_equal(r1["verdict"], exact) and not r1["from_cache"].any()
    r2 = cc.CheckCollisionBatch(robot, q)                                       # the same batch again: every row is a cache hit
    assert inner.calls == 1 and (r2["from_cache"] == 3).all()
    assert np.array_equal(r2["verdict"], exact)
    # nearby configurations (well inside the thresholds): answered from the cache, inner checker untouched
    res = np.array([robot.joints[d].resolution for d in act])
    r3 = cc.CheckCollisionBatch(robot, q + 0.01 * res / np.sqrt(len(res)))
    assert inner.calls == 1 and (r3["from_cache"] == 3).all()
    hits = cc.stats["collisionhits"] + cc.stats["freehits"]
    assert hits / float(cc.stats["checks"]) > 0.6                               # 2 of 3 passes came from the cache
    # a mixed batch: known rows from the cache, new rows in one more inner batch
    qn = sample_configs(robot, 50, seed=5, active=act)
    r4 = cc.CheckCollisionBatch(robot, np.vstack([q[:50], qn]))
    assert inner.calls == 2 and inner.configs == 350
    assert (r4["from_cache"][:50] == 3).all()
    assert np.array_equal(r4["verdict"][50:], orc.check_batch(robot, qn, active=act)["verdict"])


def test_single_state_calls_follow_the_reference_flow():
    env, robot, orc = _world()
    inner = _Inner(orc)
    cc = CacheCollisionChecker(inner, robot, collisionthresh=0.05, freespacethresh=0.05)
    act = robot.GetActiveDOFIndices()
    q = sample_configs(robot, 40, seed=6, active=act)
    exact = orc.check_batch(robot, q, active=act)["verdict"]
    from openrave_b200.environment import CollisionReport
    for rounds in range(2):
        for k in range(40):
            robot.SetDOFValues(q[k].tolist(), act)
            rep = CollisionReport()
            assert cc.CheckCollision(robot, report=rep) == bool(exact[k] & 1)
            if exact[k] & 1:
                assert rep.plink1 is not None and rep.plink1.body is robot       # the cached node carries the link pair
            assert cc.CheckStandaloneSelfCollision(robot) == bool(exact[k] & 2)
    assert cc.stats["inner_configs"] == 80 and cc.stats["checks"] == 80          # second round: all 40 + 40 from the caches
    assert cc.stats["collisionhits"] + cc.stats["freehits"] == 40
    assert cc.stats["selfcollisionhits"] + cc.stats["selffreehits"] == 40
    # other bodies bypass the cache (cachechecker.cpp:199-204)
    mug = [b for b in env.GetBodies() if b is not robot][0]
    assert cc.CheckCollision(mug) == orc.CheckCollision(mug)


import pytest


@pytest.mark.gpu
def test_gpu_cache_misses_of_a_batch_reach_the_cuda_checker_in_one_call():
    """f3: the cache in front of the CUDA checker — every miss of a batch goes to the GPU in ONE cc_check_batch call,
    repeated batches are answered by the cache, and with tight thresholds the answers are the oracle's exact ones.  The C++ form
    of the same wrapper is plugins/cudacollision/cudabatched.h::CudaCacheChecker (type-checked by tools/shim_syntax_check.sh)."""
    from openrave_b200.checker import CudaCollisionChecker
    env, robot, orc = _world()
    env2 = load_scene("wamtest1")
    robot2 = env2.GetRobots()[0]
    robot2.SetActiveDOFs(robot2.GetActiveManipulator().GetArmIndices())
    gpu = CudaCollisionChecker(env2, 0)
    env2.SetCollisionChecker(gpu)
    act = robot2.GetActiveDOFIndices()
    cc = CacheCollisionChecker(gpu, robot2, collisionthresh=0.05, freespacethresh=0.05)
    q = sample_configs(robot, 20000, seed=4, active=act)
    launches0 = gpu.GetStatistics()["launches"]
    configs0 = gpu.GetStatistics()["configs"]
    r1 = cc.CheckCollisionBatch(robot2, q)
    st = gpu.GetStatistics()
    assert len(q) <= st["configs"] - configs0 <= len(q) + 2 and 1 <= st["launches"] - launches0 <= 64     # one batched call (+ the bind-time probe and, past 16k configurations, the kernels that build the certificates)
    assert not r1["from_cache"].any()
    ref = orc.check_batch(robot, q[:1500], active=act, want_band=True)
    band = (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= 1e-4))
    assert not ((r1["verdict"][:1500] != ref["verdict"]) & ~band).any()
    r2 = cc.CheckCollisionBatch(robot2, q)                                                      # all from the cache now
    assert gpu.GetStatistics()["configs"] == st["configs"] and (r2["from_cache"] == 3).all()                # nothing reached the GPU
    assert np.array_equal(r2["verdict"], r1["verdict"])
    # single-state calls: a miss reaches the GPU's single-state overloads, a repeat does not
    qn = sample_configs(robot, 30, seed=9, active=act)
    for rounds in range(2):
        for k in range(30):
            robot2.SetDOFValues(qn[k].tolist(), act)
            cc.CheckCollision(robot2)
    assert cc.stats["inner_configs"] <= len(q) + 30
