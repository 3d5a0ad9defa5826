"""Multi-device scenes behind the C ABI (cc_scene_create_multi, include/cudacollision.h): one host call, the batch range-sharded
over the replicas, results written into disjoint slices of the caller's buffers — bit-identical to the single-device scene
(SURVEY.md 8e "verify shard-vs-single bitwise equality of outputs", V5).  On a one-GPU box the replicas share device 0, which
exercises the same mirroring and sharding code; with >= 2 devices the second test spreads them."""
import ctypes

import numpy as np
import pytest

from conftest import load_scene, sample_configs
from openrave_b200.checker import (BATCH_ALL_PAIRS, BATCH_ENV, BATCH_SELF, CudaCollisionChecker, CudaCollisionError, PinnedArray,
                                   RaveCreateCollisionChecker)
from openrave_b200 import ravemath as rm

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


def _world(devices, scene="lab1", manip=None):
    env = load_scene(scene)
    robot = env.GetRobots()[0]
    active = None
    if manip is not None:
        active = [m for m in robot.GetManipulators() if m.GetName() == manip][0].GetArmIndices()
        robot.SetActiveDOFs(active)
    cc = CudaCollisionChecker(env, devices)
    env.SetCollisionChecker(cc)
    return env, robot, cc, active


def _same(a, b):
    return (np.array_equal(a["verdict"], b["verdict"]) and np.array_equal(a["first_env_index"], b["first_env_index"])
            and np.array_equal(a["first_self_index"], b["first_self_index"]))


def _check_equal_to_single(devices, scene, manip, n):
    env1, robot1, one, active = _world(0, scene, manip)
    env2, robot2, many, _ = _world(devices, scene, manip)
    nd = ctypes.c_int(0)
    devs = (ctypes.c_int * 16)()
    assert many.lib.cc_scene_get_devices(many.scene, devs, 16, ctypes.byref(nd)) == 0 and list(devs[:nd.value]) == list(devices)
    q = sample_configs(robot1, n, seed=81, active=active)
    a = one.CheckCollisionBatch(robot1, q, active=active)
    b = many.CheckCollisionBatch(robot2, q, active=active)
    assert _same(a, b)
    assert 0.02 < (a["verdict"] != 0).mean() < 0.999
    # every replica did work: the scene-wide statistics are the sum over the replicas
    assert many.GetStatistics()["configs"] >= n and many.GetStatistics()["launches"] >= 2 * len(devices)
    # ragged sizes around the sharding threshold, complete masks, FK, distances
    for m in (1, 33, 16384 * len(devices) - 1, 16384 * len(devices) + 5):
        assert _same(one.CheckCollisionBatch(robot1, q[:m], active=active), many.CheckCollisionBatch(robot2, q[:m], active=active))
    m1 = one.CheckCollisionBatch(robot1, q[:40000], active=active, all_pairs=True)["mask"]
    m2 = many.CheckCollisionBatch(robot2, q[:40000], active=active, all_pairs=True)["mask"]
    assert np.array_equal(m1, m2)
    assert np.array_equal(one.ForwardKinematicsBatch(robot1, q[:40000], active=active), many.ForwardKinematicsBatch(robot2, q[:40000], active=active))
    d1 = one.DistanceBatch(robot1, q[:36000], active=active)
    d2 = many.DistanceBatch(robot2, q[:36000], active=active)
    assert np.array_equal(d1["distance"], d2["distance"]) and np.array_equal(d1["pair_index"], d2["pair_index"])
    # page-locked caller buffers: every replica copies its own slice
    bq = many.bind(robot2, active)
    pv, pe, ps = PinnedArray(n, np.uint32), PinnedArray(n, np.int32), PinnedArray(n, np.int32)
    qp = PinnedArray(q.shape, np.float64)
    qp.array[:] = q
    v, fe, fs, _ = bq.check(qp.array, BATCH_ENV | BATCH_SELF, out=(pv.array, pe.array, ps.array))
    assert np.array_equal(v, a["verdict"]) and np.array_equal(fe, a["first_env_index"]) and np.array_equal(fs, a["first_self_index"])
    # a moved obstacle reaches every replica (mirrored set_link_transforms + refresh)
    for e, r in ((env1, robot1), (env2, robot2)):
        mug = [x for x in e.GetBodies() if x.name == "mug1"]
        if mug:
            near = r.GetLink("wam4") or r.GetLink("r_forearm_link") or r.links[-1]
            mug[0].SetTransform(rm.t_mul(near.t, (rm.IDENT_Q, (0.0, 0.0, 0.25))))
    a2 = one.CheckCollisionBatch(robot1, q, active=active)
    b2 = many.CheckCollisionBatch(robot2, q, active=active)
    assert _same(a2, b2)
    if scene == "lab1":
        assert not _same(a, a2)
    # single-state overloads keep working on the head replica
    robot1.SetDOFValues(robot1.GetDOFValues()); robot2.SetDOFValues(robot2.GetDOFValues())
    assert env1.CheckCollision(robot1) == env2.CheckCollision(robot2)
    assert robot1.CheckSelfCollision() == robot2.CheckSelfCollision()


def test_replicas_sharing_one_device_equal_the_single_scene():
    _check_equal_to_single([0, 0, 0], "lab1", None, 120000)


def test_name_based_factory_takes_a_device_list():
    env = load_scene("lab1")
    cc = RaveCreateCollisionChecker(env, "cudacollision 0,0")
    nd = ctypes.c_int(0)
    assert cc.lib.cc_scene_get_devices(cc.scene, None, 0, ctypes.byref(nd)) == 0 and nd.value == 2
    with pytest.raises(CudaCollisionError):
        CudaCollisionChecker(env, [0, 4096])


@pytest.mark.skipif(_device_count() < 2, reason="needs >= 2 CUDA devices")
def test_shards_on_all_devices_equal_the_single_scene():
    nd = min(_device_count(), 8)
    _check_equal_to_single(list(range(nd)), "pr2test1", "rightarm_torso", 60000 * nd)
