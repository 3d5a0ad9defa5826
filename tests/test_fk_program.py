"""The FK chain program a query compiles for the device, run on the host (`cc_debug_fk_program`): device-free.

A bound query turns the body's joint list into a program of `FkStep`s (cudacollision.cu: `build_chain_steps`, `compile_fk_program`):
joints depth-first, every joint whose value the batch cannot change folded into a constant transform per link, a pose-table slot per
varying joint for `cc_fk2_kernel`, parked poses for the branches `cc_fk_kernel` returns to.  The kernels only walk that program, so
a wrong fold or slot shows up as a wrong link pose.  Here the program is walked on the host in both orders and compared with the
harness' restatement of KinBody::SetDOFValues (kinbody.cpp:2121-2485) — the same comparison the GPU tests make with the kernels'
output (tests/test_fk_ladder.py, tests/test_gpu_parity.py), for every robot of the BASELINE configurations and several active sets,
with values past the joint limits, subsets of wanted links, and a body with caller-evaluated mimic joints.
"""
import ctypes

import numpy as np
import pytest

from conftest import load_scene
from openrave_b200 import flatten
from openrave_b200.checker import load_library
from openrave_b200.kinbody import MimicExpr, external_mimic_values


def _program(robot, arr, nj, active, q, walk, wanted=None):
    lib = load_library()
    nl = len(robot.links)
    dof = np.array(robot.GetDOFValues(), dtype=np.float64)
    base = flatten.transforms8([robot.links[0].t])[0].copy()
    act = np.ascontiguousarray(active, dtype=np.int32)
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.zeros((len(q), nl, 8))
    counts = np.zeros(4, dtype=np.int32)
    w = None if wanted is None else np.ascontiguousarray(wanted, dtype=np.uint8)
    rc = lib.cc_debug_fk_program(ctypes.addressof(arr), nj, nl, robot.GetDOF(), dof.ctypes.data_as(ctypes.c_void_p),
                                 base.ctypes.data_as(ctypes.c_void_p), act.ctypes.data_as(ctypes.c_void_p), len(act),
                                 None if w is None else w.ctypes.data_as(ctypes.c_void_p), walk, q.ctypes.data_as(ctypes.c_void_p), len(q),
                                 out.ctypes.data_as(ctypes.c_void_p), counts.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0, lib.cc_last_error()
    return out, counts


def _harness(robot, full):
    q0 = robot.GetDOFValues()
    out = []
    for row in full:
        robot.SetDOFValues(row.tolist())
        out.append(flatten.transforms8(robot.GetLinkTransformations()))
    robot.SetDOFValues(q0)
    return np.array(out)


def _pose_err(a, b):
    dq = np.minimum(np.abs(a[..., :4] - b[..., :4]).max(axis=-1), np.abs(a[..., :4] + b[..., :4]).max(axis=-1))
    return max(dq.max(), np.abs(a[..., 4:7] - b[..., 4:7]).max())


def _rows(robot, active, n, rng):
    lo, hi = (np.array(x) for x in robot.GetDOFLimits())
    q = lo[active] + (hi[active] - lo[active]) * rng.random((n, len(active)))
    if len(active):
        q[::4] = lo[active] + (hi[active] - lo[active]) * (np.round(rng.random((len(q[::4]), len(active)))) * 1.4 - 0.2)   # past the limits
    full = np.tile(np.array(robot.GetDOFValues()), (n, 1))
    full[:, active] = q
    return q, full


ACTIVE_SETS = {
    "lab1": [None, [0, 1, 2, 3, 4, 5, 6], [7, 8, 9, 10], [3], []],
    "pr2test1": [None, "rightarm_torso", "leftarm", [0, 1, 2]],
    "hironxtable": [None, "rightarm", [0]],
    "wamtest1": [None, [0, 2, 4]],
}


@pytest.mark.parametrize("scene", sorted(ACTIVE_SETS))
def test_both_walks_give_the_harness_poses(scene):
    env = load_scene(scene)
    robot = env.GetRobots()[0]
    arr, nj = flatten.joint_table(robot)
    rng = np.random.default_rng(41)
    nl = len(robot.links)
    for spec in ACTIVE_SETS[scene]:
        if spec is None:
            active = list(range(robot.GetDOF()))
        elif isinstance(spec, str):
            manips = [m for m in robot.GetManipulators() if m.GetName() == spec]
            if not manips:
                continue
            active = list(manips[0].GetArmIndices())
        else:
            active = spec
        q, full = _rows(robot, active, 24, rng)
        want = _harness(robot, full)
        for walk in (0, 1):
            got, counts = _program(robot, arr, nj, active, q, walk)
            assert _pose_err(got, want) < 1e-12, (scene, spec, walk)
            # one varying step per non-static joint an active DOF drives; everything else was folded
            driven = sum(1 for j in robot.sorted_joints_all if not j.IsStatic() and
                         ((j.dofindex >= 0 and j.dofindex in active) or
                          (j.mimic is not None and j.mimic.affine() is not None and robot.GetJoint(j.mimic.affine()[0]).dofindex in active)))
            assert counts[1] == driven and counts[2] == driven + 1 and counts[0] >= driven
        # only some links wanted (what a pair list that names few links asks for): their poses are the same, the rest is dropped
        wanted = (rng.random(nl) < 0.3).astype(np.uint8)
        wanted[0] = 1
        for walk in (0, 1):
            got, counts_w = _program(robot, arr, nj, active, q, walk, wanted)
            sel = wanted.astype(bool)
            assert _pose_err(got[:, sel], want[:, sel]) < 1e-12
            assert (got[:, ~sel] == 0).all()
            assert counts_w[0] <= counts[0]


def test_caller_evaluated_mimic_columns_drive_their_joints():
    env = load_scene("lab1")
    robot = env.GetRobots()[0]
    byname = {j.name: j for j in robot.all_joints()}
    byname["JF1mimic"].mimic = MimicExpr("JF1*JF1/5 + JF2/7 + 0.3")
    byname["JF2mimic"].mimic = MimicExpr("max(JF1mimic/2, 0.25) + atan2(JF3, 1)/5")
    robot.compute_internal_information()
    ext = []
    arr, nj = flatten.joint_table(robot, externals=ext)
    rng = np.random.default_rng(43)
    for active in (list(range(robot.GetDOF())), [0, 1, 3], []):
        q, full = _rows(robot, active, 32, rng)
        clamped = full.copy()
        for j in robot.joints:
            if not j.circular:
                clamped[:, j.dofindex] = np.clip(clamped[:, j.dofindex], j.lower, j.upper)
        rows = np.hstack([q, external_mimic_values(robot, clamped, ext)])
        want = _harness(robot, full)
        for walk in (0, 1):
            got, counts = _program(robot, arr, nj, active, rows, walk)
            assert _pose_err(got, want) < 1e-12, (active, walk)
            assert counts[1] >= 2                      # the two external joints vary whatever the active set is
