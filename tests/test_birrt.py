"""BiRRT + linear shortcutting on the batched entry (csrc/cc_planner.cpp) — BASELINE.json configs[2], SURVEY.md §8 f2.

The reference planner (plugins/rplanners/rrt.h:279-659, rplanners.h:308-460, linearsmoother.cpp) checks one configuration at
a time.  The library plans with whole Extend rays / shortcuts per checker call and many problems in lockstep; these tests
pin that restructuring: (CPU) with the oracle's serial early-exit loop plugged in as the checker, plans are valid,
deterministic, and independent of batching (lockstep width, ray chunking); (GPU) the CUDA checker produces the same plans as
the oracle's serial loop, waypoint for waypoint."""
import numpy as np
import pytest

import oracle
from conftest import load_scene, sample_configs
from openrave_b200 import planning as pl
from test_edge_check import OracleBatchAdapter


def _world(scene):
    env = load_scene(scene)
    robot = env.GetRobots()[0]
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    return env, robot, orc


def _free_pairs(robot, orc, k, seed):
    act = robot.GetActiveDOFIndices()
    q = sample_configs(robot, 6 * k + 20, seed=seed, active=act)
    free = q[orc.check_batch(robot, q, active=act)["verdict"] == 0]
    assert len(free) >= 2 * k
    return free[0:2 * k:2].copy(), free[1:2 * k:2].copy()


@pytest.fixture(scope="module")
def wamtest1():
    return _world("wamtest1")


def test_edge_samples_equal_python_restatement(wamtest1):
    env, robot, orc = wamtest1
    P = pl.BirrtPlanner(robot, orc.group_checker(robot))
    c = pl.DynamicsCollisionConstraint(robot, None)
    q = sample_configs(robot, 40, seed=11, active=robot.GetActiveDOFIndices())
    for k in range(0, 40, 2):
        q1 = q[k] + 0.3 * (q[k + 1] - q[k])
        for interval in (pl.IT_Open, pl.IT_OpenStart, pl.IT_OpenEnd, pl.IT_Closed):
            a = P.EdgeSamples(q[k], q1, interval)
            b, _, _ = c.edge_samples(q[k], q1, interval)
            assert a.shape == b.shape and np.array_equal(a, b)
    assert len(P.EdgeSamples(q[0], q[0], pl.IT_Open)) == 0
    assert len(P.EdgeSamples(q[0], q[0], pl.IT_OpenStart)) == 1


def test_plans_are_valid_and_deterministic(wamtest1):
    env, robot, orc = wamtest1
    starts, goals = _free_pairs(robot, orc, 8, seed=3)
    fn, blk = orc.group_checker(robot)
    P = pl.BirrtPlanner(robot, (fn, blk))
    res = P.PlanBatch(starts, goals, seeds=np.arange(8))
    assert blk.configs_checked == P.totals["configs_serial"]        # the serial loop stops exactly where the count says
    assert P.totals["configs_submitted"] >= P.totals["configs_serial"]
    cons = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    nsmoothed = 0
    for r, s, g in zip(res, starts, goals):
        assert r["status"] == pl.PLAN_SOLVED
        for key in ("path", "path_raw"):
            p = r[key]
            assert np.array_equal(p[0], s) and np.array_equal(p[-1], g)
            assert (p >= cons.lower - 1e-12).all() and (p <= cons.upper + 1e-12).all()
        p = r["path"]
        assert r["length"] <= r["length_raw"] + 1e-9
        nsmoothed += r["length"] < r["length_raw"] - 1e-6
        # every waypoint and every edge between waypoints is collision free at the joint resolutions
        assert not orc.check_batch(robot, p)["verdict"].any()
        edges = [(p[i], p[i + 1]) for i in range(len(p) - 1)]
        assert not any(cons.CheckMany(edges, pl.IT_Open, options=0xffff))
    assert nsmoothed >= 1
    again = P.PlanBatch(starts, goals, seeds=np.arange(8))
    for a, b in zip(res, again):
        assert np.array_equal(a["path"], b["path"]) and a["iterations"] == b["iterations"]
    other = P.PlanBatch(starts[:2], goals[:2], seeds=[100, 101])
    assert any(o["iterations"] != r["iterations"] or not np.array_equal(o["path_raw"], r["path_raw"]) for o, r in zip(other, res))


def test_lockstep_and_ray_chunking_do_not_change_plans(wamtest1):
    env, robot, orc = wamtest1
    starts, goals = _free_pairs(robot, orc, 6, seed=5)
    seeds = np.arange(6) + 40
    ref = pl.BirrtPlanner(robot, orc.group_checker(robot), speculate_steps=0).PlanBatch(starts, goals, seeds)    # whole rays
    solo = pl.BirrtPlanner(robot, orc.group_checker(robot))                                                         # default: adaptive
    for k in range(6):
        r = solo.PlanPath(starts[k], goals[k], int(seeds[k]))
        assert np.array_equal(r["path"], ref[k]["path"]) and r["configs_serial"] == ref[k]["configs_serial"]
    for spec in (7, -4):                                                # fixed chunks, adaptive chunks
        chunked = pl.BirrtPlanner(robot, orc.group_checker(robot), speculate_steps=spec)
        got = chunked.PlanBatch(starts, goals, seeds)
        for a, b in zip(got, ref):
            assert np.array_equal(a["path"], b["path"]) and np.array_equal(a["path_raw"], b["path_raw"])
            assert a["configs_serial"] == b["configs_serial"] and a["configs_submitted"] <= b["configs_submitted"]
            assert a["rounds"] >= b["rounds"]


def test_shortcut_linear_equals_the_per_call_restatement(wamtest1):
    """`shortcut_linear` (ShortcutLinearPlanner, linearshortcutadvanced.cpp:76-569) as the post-processor: the library sends a whole
    shortcut (every segment of a single-DOF shortcut) to the checker at once; the Python class makes one CheckPathAllConstraints
    call per segment like the reference.  Same random stream, same raw path => the same waypoints, bit for bit."""
    env, robot, orc = wamtest1
    starts, goals = _free_pairs(robot, orc, 8, seed=3)
    P = pl.BirrtPlanner(robot, orc.group_checker(robot), post_processor="shortcut_linear")
    res = P.PlanBatch(starts, goals, seeds=np.arange(8))
    base = pl.BirrtPlanner(robot, orc.group_checker(robot)).PlanBatch(starts, goals, seeds=np.arange(8))
    cons = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    shortcuts = 0
    for k, (r, b) in enumerate(zip(res, base)):
        assert r["status"] == pl.PLAN_SOLVED and np.array_equal(r["path_raw"], b["path_raw"])    # only the post-processor differs
        S = pl.ShortcutLinearPlanner(cons, weights=P._w, step_length=0.04, max_iterations=100, seed=k)
        p, n1, n2 = S.PlanPath(r["path_raw"])
        assert np.array_equal(p, r["path"]) and n1 + n2 == r["shortcuts"]
        shortcuts += r["shortcuts"]
        assert np.array_equal(r["path"][0], starts[k]) and np.array_equal(r["path"][-1], goals[k])
        assert r["length"] <= r["length_raw"] + 1e-9
        assert not orc.check_batch(robot, r["path"])["verdict"].any()
    assert shortcuts >= 5                                             # the optimiser did fire (and one-DOF shortcuts among them)
    solo = pl.BirrtPlanner(robot, orc.group_checker(robot), post_processor="shortcut_linear", speculate_steps=0)
    for k in (4, 5):                                                  # lockstep width does not matter
        assert np.array_equal(solo.PlanPath(starts[k], goals[k], k)["path"], res[k]["path"])
    # "linearsmoother 0" (_nUseSingleDOFSmoothing = 0): shortcuts only — same raw paths, valid results, never longer than raw
    only = pl.BirrtPlanner(robot, orc.group_checker(robot), post_processor="linearsmoother 0").PlanBatch(starts, goals, seeds=np.arange(8))
    for r, b in zip(only, base):
        assert np.array_equal(r["path_raw"], b["path_raw"]) and r["length"] <= r["length_raw"] + 1e-9
        assert not orc.check_batch(robot, r["path"])["verdict"].any()


def test_colliding_endpoints_are_refused(wamtest1):
    env, robot, orc = wamtest1
    act = robot.GetActiveDOFIndices()
    q = sample_configs(robot, 200, seed=9, active=act)
    v = orc.check_batch(robot, q, active=act)["verdict"]
    bad, good = q[v != 0][0], q[v == 0]
    P = pl.BirrtPlanner(robot, orc.group_checker(robot))
    res = P.PlanBatch([bad, good[0], good[1]], [good[2], bad, good[3]], seeds=[1, 2, 3])
    assert [r["status"] for r in res] == [pl.PLAN_BAD_INITIAL, pl.PLAN_BAD_GOAL, pl.PLAN_SOLVED]
    assert len(res[0]["path"]) == 0


def test_iteration_budget_is_respected():
    """Segway-mounted WAM in lab1 with a tiny iteration budget: unsolved problems report CC_PLAN_FAILED, never a path."""
    env, robot, orc = _world("lab1")
    starts, goals = _free_pairs(robot, orc, 6, seed=21)
    P = pl.BirrtPlanner(robot, orc.group_checker(robot), max_iterations=1, smoother_iterations=0)
    res = P.PlanBatch(starts, goals, seeds=np.arange(6))
    for r in res:
        assert r["status"] in (pl.PLAN_SOLVED, pl.PLAN_FAILED)
        assert (len(r["path"]) > 0) == (r["status"] == pl.PLAN_SOLVED)
        assert r["iterations"] <= 3


@pytest.mark.gpu
def test_gpu_plans_equal_oracle_serial_plans(wamtest1):
    from openrave_b200.checker import CudaCollisionChecker
    env, robot, orc = wamtest1
    starts, goals = _free_pairs(robot, orc, 24, seed=3)
    seeds = np.arange(24)
    cpu = pl.BirrtPlanner(robot, orc.group_checker(robot)).PlanBatch(starts, goals, seeds)
    env2 = load_scene("wamtest1")
    robot2 = env2.GetRobots()[0]
    robot2.SetActiveDOFs(robot2.GetActiveManipulator().GetArmIndices())
    cc = CudaCollisionChecker(env2, 0)
    env2.SetCollisionChecker(cc)
    launches0 = cc.GetStatistics()["launches"]
    P = pl.BirrtPlanner(robot2, cc)
    gpu = P.PlanBatch(starts, goals, seeds)
    assert cc.GetStatistics()["launches"] > launches0                # the verdicts came from the CUDA kernels
    same = sum(1 for a, b in zip(cpu, gpu) if a["status"] == b["status"] and np.array_equal(a["path"], b["path"])
               and a["configs_serial"] == b["configs_serial"])
    # a plan can only differ where some sample sat inside the 1e-4 m contact band
    assert same >= 23, same
    assert all(r["status"] == pl.PLAN_SOLVED for r in gpu)
    for r in gpu:
        assert not orc.check_batch(robot, r["path"])["verdict"].any()


@pytest.mark.gpu
def test_gpu_shortcut_linear_plans_equal_oracle_serial_plans(wamtest1):
    from openrave_b200.checker import CudaCollisionChecker
    env, robot, orc = wamtest1
    starts, goals = _free_pairs(robot, orc, 16, seed=3)
    seeds = np.arange(16)
    cpu = pl.BirrtPlanner(robot, orc.group_checker(robot), post_processor="shortcut_linear").PlanBatch(starts, goals, seeds)
    env2 = load_scene("wamtest1")
    robot2 = env2.GetRobots()[0]
    robot2.SetActiveDOFs(robot2.GetActiveManipulator().GetArmIndices())
    cc = CudaCollisionChecker(env2, 0)
    env2.SetCollisionChecker(cc)
    gpu = pl.BirrtPlanner(robot2, cc, post_processor="shortcut_linear").PlanBatch(starts, goals, seeds)
    same = sum(1 for a, b in zip(cpu, gpu) if a["status"] == b["status"] and np.array_equal(a["path"], b["path"])
               and a["shortcuts"] == b["shortcuts"])
    assert same >= 15, same                                          # a difference needs a sample inside the 1e-4 m contact band
    assert sum(r["shortcuts"] for r in gpu) >= 5
    for r in gpu:
        assert r["status"] == pl.PLAN_SOLVED and not orc.check_batch(robot, r["path"])["verdict"].any()


@pytest.mark.gpu
def test_gpu_plans_pr2_right_arm_with_circular_joints():
    """PR2 right arm (two continuous joints, CO_ActiveDOFs): harder queries (tens of iterations), same plans as the serial loop."""
    from openrave_b200.checker import CO_ActiveDOFs, CudaCollisionChecker
    worlds = []
    for factory in (lambda e: oracle.OracleChecker(e, "port"), lambda e: CudaCollisionChecker(e, 0)):
        env = load_scene("pr2test1")
        robot = env.GetRobots()[0]
        arm = [m for m in robot.GetManipulators() if m.GetName() == "rightarm"][0].GetArmIndices()
        robot.SetActiveDOFs(arm)
        cc = factory(env)
        env.SetCollisionChecker(cc)
        cc.SetCollisionOptions(CO_ActiveDOFs)
        worlds.append((env, robot, cc, arm))
    (env, robot, orc, arm), (env2, robot2, cc, _) = worlds
    q = sample_configs(robot, 200, seed=3, active=arm)
    free = q[orc.check_batch(robot, q, active=arm)["verdict"] == 0]
    starts, goals, seeds = free[0:16:2].copy(), free[1:16:2].copy(), np.arange(8)
    cpu = pl.BirrtPlanner(robot, orc.group_checker(robot, arm)).PlanBatch(starts, goals, seeds)
    gpu = pl.BirrtPlanner(robot2, cc).PlanBatch(starts, goals, seeds)
    assert max(r["iterations"] for r in cpu) > 5
    same = sum(1 for a, b in zip(cpu, gpu) if a["status"] == b["status"] and np.array_equal(a["path"], b["path"]))
    assert same >= 7, same
    for r in gpu:
        assert r["status"] == pl.PLAN_SOLVED and not orc.check_batch(robot, r["path"], active=arm)["verdict"].any()


def test_basicrrt_plans_are_valid_and_batching_invariant(wamtest1):
    """BasicRrtPlanner (rrt.h:693-938) on the batched entry: a single tree with goal bias; the raw path runs from the start to a
    node within 2 x step of the goal; every edge is collision free under the oracle; lockstep width, ray chunking and the
    one-step extension (_nRRTExtentType = 1) behave as specified."""
    env, robot, orc = wamtest1
    starts, goals = _free_pairs(robot, orc, 6, seed=3)
    seeds = np.arange(6) + 7
    P = pl.BirrtPlanner(robot, orc.group_checker(robot), goal_bias=0.05, max_iterations=4000)
    res = P.PlanBatch(starts, goals, seeds, planner="basicrrt")
    cons = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    w = np.sqrt(P._w ** 2)
    for r, s, g in zip(res, starts, goals):
        assert r["status"] == pl.PLAN_SOLVED and r["nodes_backward"] == 0
        for key in ("path", "path_raw"):
            p = r[key]
            assert np.array_equal(p[0], s)
            assert np.sqrt(((w * (p[-1] - g)) ** 2).sum()) < 2 * 0.04                 # ends at the connecting node (rrt.h:820-823)
            assert not orc.check_batch(robot, p)["verdict"].any()
        p = r["path_raw"]
        assert not any(cons.CheckMany([(p[i], p[i + 1]) for i in range(len(p) - 1)], pl.IT_Open, options=0xffff))
        assert r["length"] <= r["length_raw"] + 1e-9
    solo = pl.BirrtPlanner(robot, orc.group_checker(robot), goal_bias=0.05, max_iterations=4000, speculate_steps=0)
    for k in (0, 3):
        a = solo.PlanBatch([starts[k]], [goals[k]], [seeds[k]], planner="basicrrt")[0]
        assert np.array_equal(a["path"], res[k]["path"]) and a["iterations"] == res[k]["iterations"]
    one = P.PlanBatch(starts[:3], goals[:3], seeds[:3], planner="basicrrt", extent_type=1)
    for r, full in zip(one, res):
        assert r["status"] in (pl.PLAN_SOLVED, pl.PLAN_FAILED)
        if r["status"] == pl.PLAN_SOLVED:
            assert r["iterations"] >= full["iterations"]                            # one node per iteration instead of a whole ray
            assert not orc.check_batch(robot, r["path"])["verdict"].any()
    bad = sample_configs(robot, 200, seed=9, active=robot.GetActiveDOFIndices())
    v = orc.check_batch(robot, bad, active=robot.GetActiveDOFIndices())["verdict"]
    r = P.PlanBatch([bad[v != 0][0]], [goals[0]], [1], planner="basicrrt")[0]
    assert r["status"] == pl.PLAN_BAD_INITIAL


def test_birrt_with_several_initial_and_goal_configurations(wamtest1):
    """RRTParameters::vinitialconfig / vgoalconfig with more than one configuration (rrt.h:72-88, 327-343, 443-466): one start
    and one goal per problem through the multi entry is the single entry; colliding goals / starts are skipped and keep their
    index; the plan joins one valid start to one valid goal and says which."""
    env, robot, orc = wamtest1
    act = robot.GetActiveDOFIndices()
    starts, goals = _free_pairs(robot, orc, 6, seed=3)
    seeds = np.arange(6)
    P = pl.BirrtPlanner(robot, orc.group_checker(robot))
    single = P.PlanBatch(starts, goals, seeds)
    multi = P.PlanBatchMulti([s[None] for s in starts], [g[None] for g in goals], seeds)
    for a, b in zip(single, multi):
        assert np.array_equal(a["path"], b["path"]) and a["iterations"] == b["iterations"]
        assert (b["start_index"], b["goal_index"]) == (0, 0) and (a["start_index"], a["goal_index"]) == (0, 0)
    q = sample_configs(robot, 300, seed=9, active=act)
    v = orc.check_batch(robot, q, active=act)["verdict"]
    bad, good = q[v != 0], q[v == 0]
    S = [np.array([bad[0], good[0], good[1]]), np.array([good[2]])]
    G = [np.array([bad[1], good[3], bad[2], good[4]]), np.array([bad[3], bad[4]])]
    res = P.PlanBatchMulti(S, G, [11, 12])
    r = res[0]
    assert r["status"] == pl.PLAN_SOLVED
    assert r["goal_index"] in (1, 3) and r["start_index"] in (0, 1)            # goals keep their index, starts are ranked among the valid
    assert np.array_equal(r["path"][0], S[0][1 + r["start_index"]]) and np.array_equal(r["path"][-1], G[0][r["goal_index"]])
    assert not orc.check_batch(robot, r["path"])["verdict"].any()
    assert res[1]["status"] == pl.PLAN_BAD_GOAL and res[1]["goal_index"] == -1
    assert P.PlanBatchMulti([np.array([bad[0]])], [np.array([good[0]])], [1])[0]["status"] == pl.PLAN_BAD_INITIAL


def test_birrt_minimum_goal_paths_returns_the_shortest_of_several(wamtest1):
    """RRTParameters::_minimumgoalpaths (plannerparameters.h:740; rrt.h:495-531, 540-546): with k goals wanted the planner keeps
    going after its first connection — the reached goal and everything grown from it leave the nearest-neighbour queries
    (InvalidateNodesWithParent, rplanners.h:242-266), the next goal is sampled at once — until k goals (or every valid one) have a
    path, and returns the shortest.  With 1 (the default) nothing changes.  The reference's own nearest-neighbour query treats
    invalidated nodes in a traversal-order-dependent way (rplanners.h:677-690), so which later goals connect first is "parity
    unpinned"; the properties below are what the feature promises."""
    env, robot, orc = wamtest1
    act = robot.GetActiveDOFIndices()
    q = sample_configs(robot, 400, seed=21, active=act)
    good = q[orc.check_batch(robot, q, active=act)["verdict"] == 0]
    S = [good[0][None], good[5][None]]
    G = [good[1:4], good[6:9]]
    cons = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    base = pl.BirrtPlanner(robot, orc.group_checker(robot)).PlanBatchMulti(S, G, [5, 6])
    one = pl.BirrtPlanner(robot, orc.group_checker(robot), minimum_goal_paths=1).PlanBatchMulti(S, G, [5, 6])
    for a, b in zip(base, one):
        assert a["goal_paths"] == 1 and np.array_equal(a["path"], b["path"]) and a["iterations"] == b["iterations"]
    for want in (2, 3, 7):
        P = pl.BirrtPlanner(robot, orc.group_checker(robot), minimum_goal_paths=want)
        res = P.PlanBatchMulti(S, G, [5, 6])
        again = P.PlanBatchMulti(S, G, [5, 6])
        for k, (r, r2, b) in enumerate(zip(res, again, base)):
            assert r["status"] == pl.PLAN_SOLVED and r["goal_paths"] == min(want, 3)       # never more than there are valid goals
            assert r["iterations"] >= b["iterations"]                                          # the first connection is the default run's
            assert r["length_raw"] <= b["length_raw"] + 1e-9                                  # ... and stays a candidate
            p = r["path"]
            assert np.array_equal(p[0], S[k][0]) and np.array_equal(p[-1], G[k][r["goal_index"]])
            assert not orc.check_batch(robot, p)["verdict"].any()
            edges = [(p[i], p[i + 1]) for i in range(len(p) - 1)]
            assert not any(cons.CheckMany(edges, pl.IT_Open, options=0xffff))
            assert np.array_equal(r["path"], r2["path"]) and r["goal_index"] == r2["goal_index"]


def test_birrt_goal_and_initial_samplers(wamtest1):
    """PlannerParameters::_samplegoalfn / _sampleinitialfn (rrt.h:422-437): configurations that arrive while the planner runs become
    new roots of the backward / forward tree.  A problem whose only listed goal collides is refused without a goal sampler
    ("no goals specified", rrt.h:349-355) and solved with one; the plan says which sampled goal (index after the listed ones)
    and which initial configuration it joins."""
    env, robot, orc = wamtest1
    act = robot.GetActiveDOFIndices()
    q = sample_configs(robot, 400, seed=33, active=act)
    v = orc.check_batch(robot, q, active=act)["verdict"]
    bad, good = q[v != 0], q[v == 0]
    P = pl.BirrtPlanner(robot, orc.group_checker(robot))
    S, G = [good[0][None]], [bad[0][None]]
    assert P.PlanBatchMulti(S, G, [4])[0]["status"] == pl.PLAN_BAD_GOAL
    calls = {"goal": 0, "initial": 0}

    def goal_sampler(problem):
        calls["goal"] += 1
        return good[1] if calls["goal"] == 3 else (good[2] if calls["goal"] == 5 else None)     # two goals arrive, at iterations 3 and 5

    def initial_sampler(problem):
        calls["initial"] += 1
        return good[3] if calls["initial"] == 2 else None

    P.SetSamplers(goal_sampler, initial_sampler)
    r = P.PlanBatchMulti(S, G, [4])[0]
    assert r["status"] == pl.PLAN_SOLVED and calls["goal"] >= 3 and calls["initial"] >= 2
    assert r["goal_index"] in (1, 2) and r["start_index"] in (0, 1)           # the listed goal keeps index 0 (a dummy), sampled ones follow
    goal = {1: good[1], 2: good[2]}[r["goal_index"]]
    start = {0: good[0], 1: good[3]}[r["start_index"]]
    p = r["path"]
    assert np.array_equal(p[0], start) and np.array_equal(p[-1], goal)
    cons = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    assert not orc.check_batch(robot, p)["verdict"].any()
    assert not any(cons.CheckMany([(p[i], p[i + 1]) for i in range(len(p) - 1)], pl.IT_Open, options=0xffff))
    # switched off again: the same refusal as before; and samplers that never produce anything change nothing
    P.SetSamplers(None, None)
    assert P.PlanBatchMulti(S, G, [4])[0]["status"] == pl.PLAN_BAD_GOAL
    S2, G2 = [good[0][None]], [good[1][None]]
    a = P.PlanBatchMulti(S2, G2, [9])[0]
    P.SetSamplers(lambda k: None, lambda k: None)
    b = P.PlanBatchMulti(S2, G2, [9])[0]
    assert np.array_equal(a["path"], b["path"]) and a["iterations"] == b["iterations"]


def test_birrt_may_start_without_goals_when_a_goal_sampler_is_set(wamtest1):
    """rrt.h:349-355: "no goals specified" only without _samplegoalfn."""
    env, robot, orc = wamtest1
    act = robot.GetActiveDOFIndices()
    q = sample_configs(robot, 200, seed=35, active=act)
    good = q[orc.check_batch(robot, q, active=act)["verdict"] == 0]
    P = pl.BirrtPlanner(robot, orc.group_checker(robot))
    empty = np.zeros((0, len(act)))
    with pytest.raises(RuntimeError):
        P.PlanBatchMulti([good[0][None]], [empty], [2])
    n = {"k": 0}

    def goal_sampler(problem):
        n["k"] += 1
        return good[1] if n["k"] == 4 else None

    P.SetSamplers(goal_sampler, None)
    r = P.PlanBatchMulti([good[0][None]], [empty], [2])[0]
    assert r["status"] == pl.PLAN_SOLVED and r["goal_index"] == 0 and r["iterations"] >= 4
    assert np.array_equal(r["path"][0], good[0]) and np.array_equal(r["path"][-1], good[1])
    assert not orc.check_batch(robot, r["path"])["verdict"].any()


def test_birrt_progress_callbacks_and_time_budget(wamtest1):
    """PlannerBase::_CallCallbacks at the top of every iteration (rrt.h:401-412): PA_Interrupt ends the problem with PS_Interrupted,
    PA_ReturnWithAnySolution ends a several-goals search at the first path; _nMaxPlanningTime (rrt.h:414-420) bounds the search."""
    env, robot, orc = wamtest1
    act = robot.GetActiveDOFIndices()
    q = sample_configs(robot, 400, seed=21, active=act)
    good = q[orc.check_batch(robot, q, active=act)["verdict"] == 0]
    S, G = [good[0][None]], [good[1:4]]
    base = pl.BirrtPlanner(robot, orc.group_checker(robot)).PlanBatchMulti(S, G, [5])[0]
    assert base["status"] == pl.PLAN_SOLVED
    seen = []
    P = pl.BirrtPlanner(robot, orc.group_checker(robot))
    P.SetProgressCallback(lambda problem, iteration: seen.append((problem, iteration)) or pl.PA_None)
    r = P.PlanBatchMulti(S, G, [5])[0]
    assert np.array_equal(r["path"], base["path"]) and len(seen) == r["iterations"] and seen[0] == (0, 0)
    # interrupt at the first call
    P.SetProgressCallback(lambda problem, iteration: pl.PA_Interrupt)
    ri = P.PlanBatchMulti(S, G, [5])[0]
    assert ri["status"] == pl.PLAN_INTERRUPTED and ri["iterations"] == 1 and len(ri["path"]) == 0
    # three goals wanted, "return with any solution" from the start: the search ends at the first connection, which is the default run's
    P3 = pl.BirrtPlanner(robot, orc.group_checker(robot), minimum_goal_paths=3)
    full = P3.PlanBatchMulti(S, G, [5])[0]
    P3.SetProgressCallback(lambda problem, iteration: pl.PA_ReturnWithAnySolution)
    anyr = P3.PlanBatchMulti(S, G, [5])[0]
    assert full["goal_paths"] == 3 and anyr["goal_paths"] == 1 and anyr["status"] == pl.PLAN_SOLVED
    assert np.array_equal(anyr["path_raw"], base["path_raw"]) and anyr["iterations"] <= full["iterations"]
    # a time budget that has already run out when the first iteration starts: no plan, no error
    Pt = pl.BirrtPlanner(robot, orc.group_checker(robot), max_planning_time_ms=1)
    import time
    Pt.SetProgressCallback(lambda problem, iteration: time.sleep(0.01) or pl.PA_None)
    rt = Pt.PlanBatchMulti(S, G, [5])[0]
    assert rt["status"] == pl.PLAN_FAILED and rt["iterations"] == 1
