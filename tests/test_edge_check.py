"""Edge discretisation (openrave_b200/planning.py) against the per-sample loop it replaces (SURVEY.md §7 tier V6).

CPU part (not gpu): the sample sequence obeys the reference's rules (planningutils.cpp:2692-2826, 3300-3454) and the batched
evaluation returns the same code / failing sample / time as the serial loop when both use the oracle.
GPU part: the CUDA batched entry vs the oracle's serial loop on random WAM edges."""
import numpy as np
import pytest

import oracle
from conftest import load_scene, sample_configs
from openrave_b200 import planning as pl


class OracleBatchAdapter(object):
    """Gives the oracle the product's CheckCollisionBatch signature (test-only)."""

    def __init__(self, orc):
        self.orc = orc

    def CheckCollisionBatch(self, robot, q, flags=3, active=None, all_pairs=False):
        return self.orc.check_batch(robot, q, active=active, flags=flags)


@pytest.fixture(scope="module")
def wam():
    env = load_scene("lab1")
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    robot = env.GetRobots()[0]
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
    return env, robot, orc


def test_sample_sequence_rules(wam):
    env, robot, orc = wam
    c = pl.DynamicsCollisionConstraint(robot, None)
    q0 = np.zeros(7)
    q1 = np.array([0.5, -0.3, 0.2, 0.9, 0.0, 0.1, -0.2])
    s, t, kinds = c.edge_samples(q0, q1, pl.IT_Closed)
    nsteps = int((np.abs(q1 - q0) / c.resolution + 0.99).astype(int).max())
    assert kinds[0] == 'end' and kinds[1] == 'start' and kinds.count('step') == nsteps - 1
    assert np.allclose(s[0], q1) and np.allclose(s[1], q0)
    steps = s[2:]
    assert np.allclose(steps, q0 + np.outer(np.arange(1, nsteps), (q1 - q0) / nsteps), atol=1e-12)
    assert np.allclose(t[2:], np.arange(1, nsteps) / nsteps)
    # consecutive samples never move a DOF by more than its resolution
    full = np.vstack([q0, steps, q1])
    assert (np.abs(np.diff(full, axis=0)) <= c.resolution + 1e-12).all()
    s2, t2, k2 = c.edge_samples(q0, q1, pl.IT_Open)
    assert k2.count('end') == 0 and k2.count('start') == 0 and len(s2) == nsteps - 1
    s3, _, k3 = c.edge_samples(q0, q0, pl.IT_OpenStart)
    assert k3 == ['end'] and len(s3) == 1
    s4, _, k4 = c.edge_samples(q0, q0, pl.IT_Open)
    assert len(s4) == 0


def test_batched_equals_serial_on_oracle(wam):
    env, robot, orc = wam
    cb = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    q = sample_configs(robot, 60, seed=3, active=robot.GetActiveDOFIndices())
    nfail = 0
    for k in range(0, 60, 2):
        q0, q1 = q[k], q[k] + 0.15 * (q[k + 1] - q[k])
        fa, fb = pl.ConstraintFilterReturn(), pl.ConstraintFilterReturn()
        ra = cb.Check(q0, q1, pl.IT_Closed, filterreturn=fa)
        rb = cb.CheckSerial(q0, q1, pl.IT_Closed, filterreturn=fb, env=env)
        assert ra == rb
        assert fa._invalidvalues == fb._invalidvalues and fa._fTimeWhenInvalid == fb._fTimeWhenInvalid
        assert fa._configurationtimes == fb._configurationtimes
        nfail += ra != 0
    assert 0 < nfail < 30


@pytest.mark.gpu
def test_gpu_edges_equal_oracle_serial(wam):
    from openrave_b200.checker import CudaCollisionChecker
    env, robot, orc = wam
    env2 = load_scene("lab1")
    robot2 = env2.GetRobots()[0]
    robot2.SetActiveDOFs(robot2.GetActiveManipulator().GetArmIndices())
    cc = CudaCollisionChecker(env2, 0)
    env2.SetCollisionChecker(cc)
    gpu = pl.DynamicsCollisionConstraint(robot2, cc)
    cpu = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    q = sample_configs(robot, 400, seed=8, active=robot.GetActiveDOFIndices())
    edges = [(q[k], q[k] + 0.2 * (q[k + 1] - q[k])) for k in range(0, 400, 2)]
    frs = [pl.ConstraintFilterReturn() for _ in edges]
    got = gpu.CheckMany(edges, pl.IT_Closed, filterreturns=frs)          # 200 edges, one kernel pass
    diff = 0
    for (q0, q1), g, fr in zip(edges, got, frs):
        fb = pl.ConstraintFilterReturn()
        r = cpu.Check(q0, q1, pl.IT_Closed, filterreturn=fb)
        if r != g or fr._fTimeWhenInvalid != fb._fTimeWhenInvalid:
            diff += 1
    # identical except where the first failing sample sits inside the 1e-4 m contact band
    assert diff <= 2, diff
    assert 0 < sum(1 for g in got if g) < len(got)


def test_check_edges_entry_equals_python_check(wam):
    """cc_planner_check_edges (the C entry a CheckPathAllConstraints caller switches to) against the Python restatement of
    DynamicsCollisionConstraint::Check over the same checker: return codes, failing sample, _fTimeWhenInvalid, _invalidvalues —
    linear edges with and without CFO_CheckWithPerturbation, all four intervals, and parabolic ramps."""
    env, robot, orc = wam
    act = robot.GetActiveDOFIndices()
    P = pl.BirrtPlanner(robot, orc.group_checker(robot))
    c = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    q = sample_configs(robot, 80, seed=17, active=act)
    q0 = q[0::2]
    q1 = q0 + 0.35 * (q[1::2] - q0)
    nfail = 0
    for interval in (pl.IT_Open, pl.IT_OpenStart, pl.IT_OpenEnd, pl.IT_Closed):
        for options in (0xffff, 0xffff | pl.CFO_CheckWithPerturbation):
            got = P.CheckEdges(q0, q1, interval, options, perturbation=0.1)
            for k in range(len(q0)):
                fr = pl.ConstraintFilterReturn()
                ret = c.Check(q0[k], q1[k], interval, options, fr)
                g = got[k]
                assert g["return_code"] == ret
                s, t, _ = c.edge_samples(q0[k], q1[k], interval)
                assert g["nsamples"] == len(s)
                if ret:
                    nfail += 1
                    assert np.array_equal(g["invalid_values"], np.array(fr._invalidvalues)) and g["time_when_invalid"] == fr._fTimeWhenInvalid
                    assert np.array_equal(s[g["first_invalid"]], g["invalid_values"])
                else:
                    assert g["first_invalid"] == -1
    assert nfail > 20
    # parabolic ramps (quadratic branch)
    rng = np.random.default_rng(5)
    ramps = []
    while len(ramps) < 40:
        T = rng.uniform(0.1, 1.2)
        a = c.lower + (0.2 + 0.6 * rng.random(len(act))) * (c.upper - c.lower)
        v0, v1 = rng.uniform(-2, 2, len(act)), rng.uniform(-2, 2, len(act))
        b = a + T * 0.5 * (v0 + v1)
        if (b >= c.lower).all() and (b <= c.upper).all():
            ramps.append((a, b, v0, v1, T))
    A, B, V0, V1, TT = (np.array([r[i] for r in ramps]) for i in range(5))
    got = P.CheckEdges(A, B, pl.IT_Closed, 0xffff, dq0=V0, dq1=V1, timeelapsed=TT)
    codes = set()
    for k, g in enumerate(got):
        fr = pl.ConstraintFilterReturn()
        ret = c.CheckRamp(A[k], B[k], V0[k], V1[k], TT[k], pl.IT_Closed, options=0xffff, filterreturn=fr)
        assert g["return_code"] == ret
        codes.add(ret)
        if ret and ret != pl.CFO_StateSettingError:
            assert np.array_equal(g["invalid_values"], np.array(fr._invalidvalues))
    assert len(codes) >= 2
    assert P.CheckEdges(np.zeros((0, len(act))), np.zeros((0, len(act)))) == []


@pytest.mark.gpu
def test_rebound_constraint_function_equals_the_serial_loop_and_serves_stock_planner_loops():
    """N3: the functor PlannerParameters::_checkpathvelocityconstraintsfn is rebound to
    (plugins/cudacollision/cudabatched.h::BatchedPathConstraint; planning.BatchedPathConstraint is its twin over the same C ABI call
    cc_planner_check_edges) answers every edge like DynamicsCollisionConstraint::Check's per-sample loop — return code, failing
    values, time, and the CFO_FillCheckedConfiguration samples — and an UNMODIFIED planner loop that calls it edge by edge
    (the literal ShortcutLinearPlanner) produces the plans of the library's batched shortcut."""
    from openrave_b200.checker import CudaCollisionChecker
    env = load_scene("wamtest1")
    robot = env.GetRobots()[0]
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    serial = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    env2 = load_scene("wamtest1")
    robot2 = env2.GetRobots()[0]
    robot2.SetActiveDOFs(robot2.GetActiveManipulator().GetArmIndices())
    cc = CudaCollisionChecker(env2, 0)
    env2.SetCollisionChecker(cc)
    rebound = pl.BatchedPathConstraint(robot2, cc)
    q = sample_configs(robot, 160, seed=31, active=robot.GetActiveDOFIndices())
    opts = 0xffff | pl.CFO_CheckWithPerturbation | pl.CFO_FillCheckedConfiguration
    nbad = nfail = 0
    for k in range(0, 158, 2):
        a, b = q[k], q[k] + 0.35 * (q[k + 1] - q[k])
        for interval in (pl.IT_Closed, pl.IT_OpenStart, pl.IT_Open):
            f1, f2 = pl.ConstraintFilterReturn(), pl.ConstraintFilterReturn()
            r1 = serial.CheckSerial(a, b, interval, opts, f1, env=env)
            r2 = rebound.Check(a, b, interval, opts, f2)
            same = (r1 == r2 and f1._returncode == f2._returncode and np.allclose(f1._invalidvalues, f2._invalidvalues, atol=0)
                    and f1._fTimeWhenInvalid == f2._fTimeWhenInvalid and len(f1._configurations) == len(f2._configurations)
                    and all(np.array_equal(x, y) for x, y in zip(f1._configurations, f2._configurations))
                    and f1._configurationtimes == f2._configurationtimes)
            nbad += 0 if same else 1
            nfail += 1 if r1 else 0
    assert nfail > 20 and nbad <= 2, (nfail, nbad)        # a difference needs a sample inside the 1e-4 m contact band
    assert rebound.edges == 79 * 3 and rebound.samples > rebound.edges
    # a stock-style planner loop over the rebound function == the library's batched shortcut_linear, waypoint for waypoint
    free = q[cc.CheckCollisionBatch(robot2, q)["verdict"] == 0]
    starts, goals = free[0:12:2].copy(), free[1:12:2].copy()
    P = pl.BirrtPlanner(robot2, cc, post_processor="shortcut_linear")
    res = P.PlanBatch(starts, goals, seeds=np.arange(len(starts)))
    same = 0
    for k, r in enumerate(res):
        assert r["status"] == pl.PLAN_SOLVED
        S = pl.ShortcutLinearPlanner(rebound, weights=P._w, step_length=0.04, max_iterations=100, seed=k)
        p, n1, n2 = S.PlanPath(r["path_raw"])
        same += int(np.array_equal(p, r["path"]) and n1 + n2 == r["shortcuts"])
    assert same == len(res), same
