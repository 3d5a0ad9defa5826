"""Parabolic-ramp edges: the quadratic branch of `DynamicsCollisionConstraint::Check` (src/libopenrave/planningutils.cpp:2660-2681,
2726-2766, 2844-3319), taken when the caller passes end-point velocities and timeelapsed > 0 (the parabolic smoothers'
feasibility checks).  The sample sequence does not depend on any collision result, so — like the linear branch — the whole ramp
is generated up front (`cc_planner_ramp_samples`, csrc/cc_planner.cpp) and checked in one batch.

Pins: the C++ sampler == the scalar Python restatement (`planning.DynamicsCollisionConstraint.ramp_samples`) bit for bit; every
reported sample lies on q0 + dq0 t + a t^2/2; consecutive samples are never farther apart than 1.01 x the joint resolutions
(the property the reference's time-scaling loop exists for); the walk ends at q1; `CheckRamp` returns the code of the first
failing sample in the reference's order.  No OpenRAVE runs here, so the exact sequence is "parity unpinned" against a live
reference."""
import numpy as np
import pytest

import oracle
from conftest import load_scene, sample_configs
from openrave_b200 import planning as pl
from test_edge_check import OracleBatchAdapter

INTERVALS = (pl.IT_Open, pl.IT_OpenStart, pl.IT_OpenEnd, pl.IT_Closed)


@pytest.fixture(scope="module")
def wam():
    env = load_scene("wamtest1")
    robot = env.GetRobots()[0]
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    return env, robot, orc


def _ramps(c, count, seed, vmax=1.0):
    """Consistent ramps inside the joint limits: q1 = q0 + T (dq0 + dq1) / 2."""
    rng = np.random.default_rng(seed)
    n = len(c.active)
    out = []
    while len(out) < count:
        T = rng.uniform(0.05, 1.5)
        q0 = c.lower + (0.15 + 0.7 * rng.random(n)) * (c.upper - c.lower)
        dq0 = rng.uniform(-vmax, vmax, n) * rng.choice([0, 1, 1, 1], n)
        dq1 = rng.uniform(-vmax, vmax, n) * rng.choice([0, 1, 1, 1], n)
        if len(out) % 5 == 0:
            dq1 = dq0.copy()                                            # constant velocity: the a == 0 path (:2940-2944)
        q1 = q0 + T * 0.5 * (dq0 + dq1)
        if (q1 < c.lower).any() or (q1 > c.upper).any():
            continue
        out.append((q0, q1, dq0, dq1, T))
    return out


def test_ramp_samples_equal_python_restatement_and_lie_on_the_parabola(wam):
    env, robot, orc = wam
    P = pl.BirrtPlanner(robot, orc.group_checker(robot))
    c = pl.DynamicsCollisionConstraint(robot, None)
    nerr = ntot = 0
    for q0, q1, dq0, dq1, T in _ramps(c, 120, seed=7):
        for interval in INTERVALS:
            a = P.RampSamples(q0, q1, dq0, dq1, T, interval)
            b = c.ramp_samples(q0, q1, dq0, dq1, T, interval)
            ntot += 1
            if a is None or b is None:
                assert a is None and b is None                          # CFO_StateSettingError in both
                nerr += 1
                continue
            s, t, f = a
            assert s.shape == b[0].shape and np.array_equal(s, b[0])
            assert np.array_equal(np.nan_to_num(t, nan=-1.0), np.nan_to_num(np.array(b[1]), nan=-1.0)) and list(f) == list(b[2])
            # order: q1 first when the interval is closed at the end, then q0 when closed at the start, then increasing times
            k = 0
            if interval in (pl.IT_OpenStart, pl.IT_Closed):
                assert np.array_equal(s[0], q1) and t[0] == T
                k = 1
            if interval in (pl.IT_OpenEnd, pl.IT_Closed):
                assert np.array_equal(s[k], q0) and t[k] == 0.0
                k += 1
            walk_t, walk, kind = t[k:], s[k:], np.asarray(f[k:])
            # linear sub-steps (time NaN, checked but not reported) appear where a time-scaled step still exceeds 1.01 x a
            # resolution (:3163-3196), and as the tail between the last ramp sample and q1 (:3207-3258, checked before it)
            sub = np.isnan(walk_t)
            assert np.array_equal(sub, (kind == 0) | (kind == 2)) and (kind[~sub] == 1).all()
            ramp_t, ramp = walk_t[~sub], walk[~sub]
            assert (np.diff(ramp_t) > 0).all() and (ramp_t > 0).all() and (ramp_t <= T + 1e-12).all()
            acc = (dq1 - dq0) / T
            para = q0[None, :] + ramp_t[:, None] * dq0[None, :] + 0.5 * ramp_t[:, None] ** 2 * acc[None, :]
            assert len(ramp) == 0 or np.abs(para - ramp).max() < 1e-12
            # resolution property in time order: q0 -> (sub-steps ->) ramp samples -> tail sub-steps -> q1
            tail = kind == 2
            chain = np.vstack([q0[None, :], walk[~tail], walk[tail], q1[None, :]])
            assert (np.abs(np.diff(chain, axis=0)) <= 1.01 * c.resolution[None, :] + 1e-9).all()
    assert ntot == 480 and nerr < 0.1 * ntot                            # the "fStep did not reach fLargestStep" cases (:3198-3205)


def test_constant_velocity_ramp_visits_the_linear_edge(wam):
    """dq0 == dq1: the ramp is the straight segment, walked in equal increments of the DOF that needs the most steps
    (timestep = fBestNewStep / velocity, :2940-2944): nsteps - 1 evenly spaced interior points."""
    env, robot, orc = wam
    P = pl.BirrtPlanner(robot, orc.group_checker(robot))
    c = pl.DynamicsCollisionConstraint(robot, None)
    q0 = c.lower + 0.4 * (c.upper - c.lower)
    v = np.array([0.3, -0.2, 0.1, 0.25, -0.15, 0.05, 0.2])
    T = 0.8
    q1 = q0 + T * v
    s, t, f = P.RampSamples(q0, q1, v, v, T, pl.IT_Open)
    nsteps = int((np.abs(q1 - q0) / c.resolution + 0.99).astype(int).max())
    assert len(s) == nsteps - 1 and (f == 1).all()
    want = q0[None, :] + (np.arange(1, nsteps)[:, None] / nsteps) * (q1 - q0)[None, :]
    assert np.abs(s - want).max() < 1e-12 and np.abs(t - T * np.arange(1, nsteps) / nsteps).max() < 1e-12
    # closed interval: q1 first, then q0, then the same walk
    s2, t2, f2 = P.RampSamples(q0, q1, v, v, T, pl.IT_Closed)
    assert np.array_equal(s2[0], q1) and np.array_equal(s2[1], q0) and np.array_equal(s2[2:], s)


def test_check_ramp_reports_the_first_failing_sample(wam):
    env, robot, orc = wam
    adapter = OracleBatchAdapter(orc)
    c = pl.DynamicsCollisionConstraint(robot, adapter)
    act = robot.GetActiveDOFIndices()
    nfail = nfree = 0
    for q0, q1, dq0, dq1, T in _ramps(c, 60, seed=21, vmax=2.0):
        r = c.ramp_samples(q0, q1, dq0, dq1, T, pl.IT_Closed)
        fr = pl.ConstraintFilterReturn()
        ret = c.CheckRamp(q0, q1, dq0, dq1, T, pl.IT_Closed, options=0xffff | pl.CFO_FillCheckedConfiguration, filterreturn=fr)
        if r is None:
            assert ret == pl.CFO_StateSettingError
            continue
        s, t, f = r
        v = orc.check_batch(robot, s, active=act)["verdict"]
        bad = np.nonzero(v)[0]
        if len(bad) == 0:
            nfree += 1
            assert ret == 0
            # reported: q0, the walk, q1 last (:3611-3623)
            got = np.array(fr._configurations)
            rep = np.asarray(f[1:]) == 1
            want = np.vstack([s[1:][rep], s[:1]]) if f[0] == 1 else s[1:][rep]
            assert np.array_equal(got, want)
        else:
            nfail += 1
            k = int(bad[0])
            assert ret == (pl.CFO_CheckEnvCollisions if v[k] & 1 else pl.CFO_CheckSelfCollisions)
            assert np.array_equal(np.array(fr._invalidvalues), s[k])
    assert nfail >= 3 and nfree >= 3


@pytest.mark.gpu
def test_gpu_check_ramp_equals_oracle(wam):
    from openrave_b200.checker import CudaCollisionChecker
    env, robot, orc = wam
    env2 = load_scene("wamtest1")
    robot2 = env2.GetRobots()[0]
    robot2.SetActiveDOFs(robot2.GetActiveManipulator().GetArmIndices())
    cc = CudaCollisionChecker(env2, 0)
    env2.SetCollisionChecker(cc)
    cpu = pl.DynamicsCollisionConstraint(robot, OracleBatchAdapter(orc))
    gpu = pl.DynamicsCollisionConstraint(robot2, cc)
    same = total = 0
    for q0, q1, dq0, dq1, T in _ramps(cpu, 60, seed=21, vmax=2.0):
        fa, fb = pl.ConstraintFilterReturn(), pl.ConstraintFilterReturn()
        a = cpu.CheckRamp(q0, q1, dq0, dq1, T, pl.IT_Closed, filterreturn=fa)
        b = gpu.CheckRamp(q0, q1, dq0, dq1, T, pl.IT_Closed, filterreturn=fb)
        total += 1
        same += (a == b and fa._invalidvalues == fb._invalidvalues and fa._configurations == fb._configurations)
    assert same >= total - 1, (same, total)                              # a difference needs a sample inside the contact band
