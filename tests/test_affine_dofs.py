"""Affine active DOFs (RobotBase::SetActiveDOFValues with X | Y | Z | RotationAxis | Rotation3D | RotationQuat,
src/libopenrave/robot.cpp:1193-1225, RaveGetTransformFromAffineDOFValues libopenrave.cpp:1868-1924) — base planning.
CPU: the harness robot and the oracle agree on every link pose.  GPU: batched FK and verdicts of a PR2 base-planning batch
against the oracle."""
import numpy as np
import pytest

import oracle
from conftest import ORACLE_KIND, load_scene
from openrave_b200 import flatten

DOF_X, DOF_Y, DOF_Z, DOF_RotationAxis, DOF_Rotation3D, DOF_RotationQuat = 1, 2, 4, 8, 16, 32


def _pr2(affine, axis=(0.0, 0.0, 1.0)):
    env = load_scene("pr2test1")
    robot = env.GetRobots()[0]
    arm = [m for m in robot.GetManipulators() if m.GetName() == "rightarm"][0].GetArmIndices()
    robot.SetActiveDOFs(arm, affine, axis)
    return env, robot, list(arm)


def _sample(robot, arm, affine, n, seed, spread=0.6):
    rng = np.random.default_rng(seed)
    lo, hi = robot.GetDOFLimits()
    lo = np.array([lo[i] for i in arm]); hi = np.array([hi[i] for i in arm])
    cols = [lo + rng.random((n, len(arm))) * (hi - lo)]
    base = np.array(robot.links[0].t[1])
    for bit, k in ((DOF_X, 0), (DOF_Y, 1), (DOF_Z, 2)):
        if affine & bit:
            cols.append(base[k] + ((rng.random((n, 1)) * 0.08 + 0.03) if bit == DOF_Z else (rng.random((n, 1)) - 0.5) * 2 * spread))
    if affine & DOF_RotationAxis:
        cols.append((rng.random((n, 1)) - 0.5) * 2 * np.pi)
    elif affine & DOF_Rotation3D:
        cols.append((rng.random((n, 3)) - 0.5) * np.array([0.06, 0.06, 4.0]))       # small tilts, any yaw (axis * angle)
    elif affine & DOF_RotationQuat:
        cols.append(rng.normal(size=(n, 4)) * np.array([1.0, 0.02, 0.02, 1.0]))     # mostly yaw; not normalised on purpose: the reference normalises
    return np.hstack(cols)


@pytest.mark.parametrize("affine,axis", [(DOF_X | DOF_Y | DOF_RotationAxis, (0.0, 0.0, 1.0)), (DOF_X | DOF_Y | DOF_Z | DOF_Rotation3D, None),
                                         (DOF_Y | DOF_RotationQuat, None), (DOF_RotationAxis, (0.0, 0.6, 0.8))])
def test_harness_and_oracle_agree_on_affine_forward_kinematics(affine, axis):
    env, robot, arm = _pr2(affine, axis or (0.0, 0.0, 1.0))
    assert robot.GetActiveDOF() == len(arm) + flatten.affine_count(affine)
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    q = _sample(robot, arm, affine, 12, seed=5)
    fk = orc.check_batch(robot, q, flags=0, want_fk=True)["fk"]
    t0 = robot.GetLinkTransformations()
    for k in range(len(q)):
        robot.SetActiveDOFValues(q[k].tolist())
        mine = flatten.transforms8(robot.GetLinkTransformations())
        dq = np.minimum(np.abs(mine[:, :4] - fk[k][:, :4]).max(), np.abs(mine[:, :4] + fk[k][:, :4]).max())
        assert dq < 1e-12 and np.abs(mine[:, 4:7] - fk[k][:, 4:7]).max() < 1e-12
        robot.SetLinkTransformations(t0)                                       # the oracle's base is the pose at sync time
    # the affine components that are NOT active keep the bound pose: Z here, when only X | Y move
    if not affine & DOF_Z:
        assert np.allclose(fk[:, 0, 6], t0[0][1][2], atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("affine", [DOF_X | DOF_Y | DOF_RotationAxis, DOF_X | DOF_Y | DOF_Z | DOF_Rotation3D, DOF_X | DOF_RotationQuat])
def test_gpu_base_planning_batch_equals_oracle(affine):
    from openrave_b200.checker import CO_ActiveDOFs, CudaCollisionChecker
    env, robot, arm = _pr2(affine)
    env2, robot2, _ = _pr2(affine)
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    orc = oracle.OracleChecker(env2, ORACLE_KIND)
    env2.SetCollisionChecker(orc)
    n = 30000
    q = _sample(robot, arm, affine, n, seed=7)
    got_fk = cc.ForwardKinematicsBatch(robot, q[:400])
    ref_fk = orc.check_batch(robot2, q[:400], flags=0, want_fk=True)["fk"]
    dq = np.minimum(np.abs(got_fk[..., :4] - ref_fk[..., :4]).max(-1), np.abs(got_fk[..., :4] + ref_fk[..., :4]).max(-1)).max()
    assert dq < 1e-12 and np.abs(got_fk[..., 4:7] - ref_fk[..., 4:7]).max() < 1e-11
    for opts in (0, CO_ActiveDOFs):
        cc.SetCollisionOptions(opts)
        orc.SetCollisionOptions(opts)
        got = cc.CheckCollisionBatch(robot, q)
        m = 1500
        ref = orc.check_batch(robot2, q[:m], want_band=True)
        band = (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= 1e-4))
        bad = got["verdict"][:m] != ref["verdict"]
        assert not (bad & ~band).any(), (opts, int((bad & ~band).sum()))
        ok = ~band
        fe = got["first_env"][:m].copy()
        hit = fe[:, 0] >= 0
        fe[hit, 1] = [orc._ids[id([b for b in env2.GetBodies() if b.name == cc.body_from_id(int(x)).name][0])] for x in fe[hit, 1]]
        assert (fe[ok] == ref["first_env"][ok]).all() and (got["first_self"][:m][ok] == ref["first_self"][ok]).all()
        assert 0.05 < (ref["verdict"] != 0).mean() < 0.995
        # every env pair depends on the base pose: nothing involving the environment may be hoisted; self pairs still are
        info = got["query"].info
        if opts == 0:                       # (with CO_ActiveDOFs only self pairs an active joint can change are listed at all)
            assert info.npairs_hoisted > 0
            quads = got["query"].pairs
            assert (quads[:info.npairs_env, 3] == 0).all()
    # pose mode agrees with FK mode (the single-state overloads read current transforms)
    bq = cc.bind(robot)
    v1, fe1, fs1, _ = bq.check(q[:300])
    v2, fe2, fs2, _ = bq.check_poses(bq.fk(q[:300]))
    assert np.array_equal(v1, v2) and np.array_equal(fe1, fe2) and np.array_equal(fs1, fs2)
