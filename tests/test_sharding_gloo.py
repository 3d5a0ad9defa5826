"""world_size-2 gloo test of the multi-GPU host logic (openrave_b200/sharding.py): range shards cover the batch exactly
once and the gathered results equal the single-rank results (SURVEY.md §7 tier V5).  The compute function is a stand-in
for the CUDA call — the sharding layer never looks inside it.  not gpu."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_check(q):
    """Deterministic per-configuration 'verdict / first pair' words."""
    s = (np.abs(q).sum(axis=1) * 1000).astype(np.int64)
    verdict = (s % 4).astype(np.uint32)
    fe = np.where(verdict & 1, s % 97, -1).astype(np.int32)
    fs = np.where(verdict & 2, 100 + s % 31, -1).astype(np.int32)
    return verdict, fe, fs


def _worker(rank, world, port, n, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from openrave_b200 import sharding
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    q = np.random.default_rng(5).random((n, 7))
    v, fe, fs = sharding.check_batch_sharded(fake_check, q, rank, world)
    np.savez(out % rank, v=v, fe=fe, fs=fs)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    from openrave_b200 import sharding
    for n in (0, 1, 7, 100, 100003):
        for world in (1, 2, 4, 8):
            edges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = sharding.shard_sizes(n, world)
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n", [1001, 3])
def test_two_rank_gather_equals_single(tmp_path, n):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r%d.npz")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    q = np.random.default_rng(5).random((n, 7))
    v, fe, fs = fake_check(q)
    for r in range(2):
        d = np.load(out % r)
        assert (d["v"] == v).all() and (d["fe"] == fe).all() and (d["fs"] == fs).all()


def _plan_worker(rank, world, port, k, out):
    """Planning problems shard like configurations: contiguous ranges, one planner per rank, statuses and path lengths gathered."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle
    from conftest import load_scene, sample_configs
    from openrave_b200 import planning as pl
    from openrave_b200 import sharding
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    env = load_scene("wamtest1")
    robot = env.GetRobots()[0]
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    q = sample_configs(robot, 8 * k + 20, seed=3, active=robot.GetActiveDOFIndices())
    free = q[orc.check_batch(robot, q)["verdict"] == 0]
    starts, goals, seeds = free[0:2 * k:2], free[1:2 * k:2], np.arange(k)
    lo, hi = sharding.shard_range(k, rank, world)
    res = pl.BirrtPlanner(robot, orc.group_checker(robot)).PlanBatch(starts[lo:hi], goals[lo:hi], seeds[lo:hi])
    mine = torch.zeros(k, 2, dtype=torch.float64)
    for i, r in enumerate(res):
        mine[lo + i, 0] = r["status"]
        mine[lo + i, 1] = r["length"]
    dist.all_reduce(mine)                       # disjoint rows: the sum is the gather
    if rank == 0:
        np.save(out, mine.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_planning_equals_single(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    from conftest import load_scene, sample_configs
    from openrave_b200 import planning as pl
    k = 6
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "plans.npy")
    mp.spawn(_plan_worker, args=(2, port, k, out), nprocs=2, join=True)
    got = np.load(out)
    env = load_scene("wamtest1")
    robot = env.GetRobots()[0]
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    q = sample_configs(robot, 8 * k + 20, seed=3, active=robot.GetActiveDOFIndices())
    free = q[orc.check_batch(robot, q)["verdict"] == 0]
    res = pl.BirrtPlanner(robot, orc.group_checker(robot)).PlanBatch(free[0:2 * k:2], free[1:2 * k:2], np.arange(k))
    assert np.array_equal(got[:, 0], [r["status"] for r in res])
    assert np.allclose(got[:, 1], [r["length"] for r in res], rtol=0, atol=0)
