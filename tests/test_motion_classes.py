"""What a bound query evaluates once at bind time instead of per configuration (cc_query_info.npairs_hoisted), device-free.

The reference evaluates every link pair at every configuration (collisionPQP.h:364-417, 432-498).  The library skips pairs whose
verdict cannot change during a batch: environment pairs of links that do not move, self pairs whose relative pose is fixed.  Both
classifications are host code (`MotionClasses` in cudacollision.cu) exported as `cc_debug_motion_classes`; here they are checked
against brute force — the harness' SetDOFValues restatement at random values of the active DOFs — on the robots of the BASELINE
configurations, and on hand-built chains that contain the case the first form of the rule got wrong.
"""
import ctypes
import os

import numpy as np
import pytest

import oracle
from conftest import load_scene
from openrave_b200 import flatten, ravemath as rm
from openrave_b200.checker import load_library
from openrave_b200.flatten import CC_MIMIC_EXTERNAL, cc_joint


def _classes(arr, nj, nl, ndof, affects, active, pairs, base_moves=0):
    lib = load_library()
    affects = np.ascontiguousarray(affects, dtype=np.uint8)
    act = np.ascontiguousarray(active, dtype=np.int32)
    pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    moves = np.zeros(nl, dtype=np.uint8)
    inv = np.zeros(max(len(pr), 1), dtype=np.uint8)
    rc = lib.cc_debug_motion_classes(ctypes.addressof(arr), nj, nl, ndof, affects.ctypes.data_as(ctypes.c_void_p),
                                     act.ctypes.data_as(ctypes.c_void_p), len(act), base_moves, pr.ctypes.data_as(ctypes.c_void_p),
                                     len(pr), moves.ctypes.data_as(ctypes.c_void_p), inv.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return moves.astype(bool), inv[:len(pr)].astype(bool)


def _same(t1, t2, tol=1e-9):
    q1, p1 = np.array(t1[0]), np.array(t1[1])
    q2, p2 = np.array(t2[0]), np.array(t2[1])
    return min(np.abs(q1 - q2).max(), np.abs(q1 + q2).max()) < tol and np.abs(p1 - p2).max() < tol


ACTIVE_SETS = {
    "lab1": [None, [0, 1, 2, 3, 4, 5, 6], [7, 8, 9, 10], [7], [3]],
    "pr2test1": [None, "rightarm_torso", "leftarm", [0, 1, 2]],
    "hironxtable": [None, "rightarm", [0]],
}


@pytest.mark.parametrize("scene", sorted(ACTIVE_SETS))
def test_classes_against_brute_force_fk(scene):
    env = load_scene(scene)
    robot = env.GetRobots()[0]
    ndof, nl = robot.GetDOF(), len(robot.links)
    arr, nj = flatten.joint_table(robot)
    aff = flatten.affects_table(robot)
    lo, hi = (np.array(x) for x in robot.GetDOFLimits())
    q0 = np.array(robot.GetDOFValues())
    pairs = [(a, b) for a in range(nl) for b in range(a + 1, nl)]
    env.SetCollisionChecker(oracle.OracleChecker(env, "port"))            # GetNonAdjacentLinks probes the load pose
    nonadjacent = {(v & 0xffff, v >> 16) for v in robot.GetNonAdjacentLinks(0)}
    rng = np.random.default_rng(23)
    for spec in ACTIVE_SETS[scene]:
        if spec is None:
            active = list(range(ndof))
        elif isinstance(spec, str):
            manips = [m for m in robot.GetManipulators() if m.GetName() == spec]
            if not manips:
                continue
            active = list(manips[0].GetArmIndices())
        else:
            active = spec
        moves, inv = _classes(arr, nj, nl, ndof, aff, active, pairs)
        poses = []
        for _ in range(5):
            q = q0.copy()
            q[active] = lo[active] + rng.random(len(active)) * (hi[active] - lo[active])
            robot.SetDOFValues(q.tolist())
            poses.append(robot.GetLinkTransformations())
        robot.SetDOFValues(q0.tolist())
        # a link that is said not to move never does; one that is said to move does (every active DOF here has a non-empty range)
        for l in range(nl):
            still = all(_same(poses[0][l], p[l]) for p in poses[1:])
            assert still == (not moves[l]), (scene, spec, l)
        # the same for the relative pose of every pair of links.  The first form of the rule (the two links have the same set of
        # affecting active DOFs) is evaluated next to it: it calls the Barrett hand's Finger0-1 / Finger0-2 (a mimic joint between
        # them follows JF1, which moves both), its two spread links (one joint follows the other's DOF) and the PR2's finger links
        # (all four follow one gripper DOF) fixed — wrongly.  None of those pairs is in GetNonAdjacentLinks (adjacent, or touching at
        # the load pose like the PR2's two finger tips), so no pair list of these robots ever held one; a robot loaded with its
        # gripper open would have
        rel = [[rm.t_inv(p[l]) for l in range(nl)] for p in poses]
        for k, (a, b) in enumerate(pairs):
            fixed = all(_same(rm.t_mul(rel[0][a], poses[0][b]), rm.t_mul(rel[i][a], poses[i][b])) for i in range(1, len(poses)))
            assert fixed == bool(inv[k]), (scene, spec, a, b)
            old_rule = all(aff[d, a] == aff[d, b] for d in active)
            if old_rule != bool(inv[k]):
                assert old_rule and not fixed and (a, b) not in nonadjacent, (scene, spec, a, b)


def _chain(specs):
    """specs: per joint (parent, child, dof_index, mimic_dof, is_static[, reserved]); revolute about z, identity frames."""
    arr = (cc_joint * len(specs))()
    for k, sp in enumerate(specs):
        c = arr[k]
        c.type = 1
        c.parent_link, c.child_link, c.dof_index, c.mimic_dof, c.is_static = sp[:5]
        c.reserved = sp[5] if len(sp) > 5 else 0
        c.mimic_a, c.mimic_b = 1.0, 0.0
        c.lower, c.upper = -1.0, 1.0
        c.left[0] = c.right[0] = 1.0
        c.axis[2] = 1.0
    return arr, len(specs)


def test_a_mimic_joint_between_two_links_that_follows_a_dof_above_both():
    # 0 -(dof 0)-> 1 -(mimic of dof 0)-> 2 -(static)-> 3, and a side branch 0 -(dof 1)-> 4 -(static)-> 5
    arr, nj = _chain([(0, 1, 0, -1, 0), (1, 2, -1, 0, 0), (2, 3, -1, -1, 1), (0, 4, 1, -1, 0), (4, 5, -1, -1, 1)])
    aff = np.array([[0, 1, 1, 1, 0, 0], [0, 0, 0, 0, 1, 1]], dtype=np.uint8)        # DoesAffect follows the mimic equation
    pairs = [(1, 3), (1, 2), (2, 3), (0, 3), (4, 5), (3, 5), (0, 5)]
    moves, inv = _classes(arr, nj, 6, 2, aff, [0, 1], pairs)
    assert moves.tolist() == [False, True, True, True, True, True]
    # (1, 3): both links are affected by DOF 0 and by nothing else, yet the mimic joint between them turns with it
    assert inv.tolist() == [False, False, True, False, True, False, False]
    # only DOF 1 active: everything on the first branch is fixed, in the world and relative to each other
    moves, inv = _classes(arr, nj, 6, 2, aff, [1], pairs)
    assert moves.tolist() == [False, False, False, False, True, True]
    assert inv.tolist() == [True, True, True, True, True, False, False]
    # affine DOFs active: every link moves, relative poses are as before
    moves, inv2 = _classes(arr, nj, 6, 2, aff, [1], pairs, base_moves=1)
    assert moves.all() and inv2.tolist() == inv.tolist()


def test_links_below_a_caller_evaluated_mimic_joint_always_move():
    # 0 -(dof 0)-> 1 -(external mimic, column 0)-> 2 -(static)-> 3
    arr, nj = _chain([(0, 1, 0, -1, 0), (1, 2, -1, CC_MIMIC_EXTERNAL, 0, 0), (2, 3, -1, -1, 1)])
    aff = np.array([[0, 1, 1, 1]], dtype=np.uint8)
    pairs = [(0, 1), (1, 2), (1, 3), (2, 3), (0, 3)]
    moves, inv = _classes(arr, nj, 4, 1, aff, [], pairs)               # no DOF active at all
    assert moves.tolist() == [False, False, True, True]
    assert inv.tolist() == [True, False, False, True, False]
    moves, inv = _classes(arr, nj, 4, 1, aff, [0], pairs)
    assert moves.tolist() == [False, True, True, True]
    assert inv.tolist() == [False, False, False, True, False]


def test_bad_arguments_are_refused():
    lib = load_library()
    arr, nj = _chain([(0, 1, 0, -1, 0)])
    out = np.zeros(2, dtype=np.uint8)
    aff = np.ones((1, 2), dtype=np.uint8)
    bad_active = np.array([3], dtype=np.int32)
    assert lib.cc_debug_motion_classes(ctypes.addressof(arr), nj, 2, 1, aff.ctypes.data_as(ctypes.c_void_p),
                                       bad_active.ctypes.data_as(ctypes.c_void_p), 1, 0, None, 0, out.ctypes.data_as(ctypes.c_void_p), None) != 0
    bad_pair = np.array([0, 7], dtype=np.int32)
    act = np.array([0], dtype=np.int32)
    assert lib.cc_debug_motion_classes(ctypes.addressof(arr), nj, 2, 1, aff.ctypes.data_as(ctypes.c_void_p), act.ctypes.data_as(ctypes.c_void_p),
                                       1, 0, bad_pair.ctypes.data_as(ctypes.c_void_p), 1, out.ctypes.data_as(ctypes.c_void_p),
                                       out.ctypes.data_as(ctypes.c_void_p)) != 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/robots"), reason="needs /root/reference")
def test_soundness_on_every_shipped_robot():
    """The two flags may only be set when nothing can change (the reference evaluates every pair at every configuration): for
    every robot file under src/robots and src/collada_robots that loads, with all DOFs active and with a random half of them, no
    link flagged still moves and no pair flagged invariant changes its relative pose.  (pumabarrett.robot.xml — a hand that no joint
    ties to its arm — is left out: the query builder refuses that forest before these flags are used.)"""
    import glob
    from openrave_b200.environment import Environment
    files = sorted(glob.glob("/root/reference/src/robots/*") + glob.glob("/root/reference/src/collada_robots/*.zae"))
    rng = np.random.default_rng(2)
    bodies = 0
    for f in files:
        if not f.endswith((".xml", ".zae")) or f.endswith("pumabarrett.robot.xml"):
            continue
        env = Environment()
        try:
            env.Load(f)
        except (NotImplementedError, IOError):
            continue
        for b in env.GetBodies():
            if b.GetDOF() == 0 or len(b.links) < 2 or len(b.links) > 50:
                continue
            arr, nj = flatten.joint_table(b, externals=[])
            aff = flatten.affects_table(b)
            nl, ndof = len(b.links), b.GetDOF()
            lo, hi = (np.array(x) for x in b.GetDOFLimits())
            q0 = np.array(b.GetDOFValues())
            pairs = [(a, c) for a in range(nl) for c in range(a + 1, nl)]
            for active in (list(range(ndof)), sorted(rng.choice(ndof, max(1, ndof // 2), replace=False).tolist())):
                moves, inv = _classes(arr, nj, nl, ndof, aff, active, pairs)
                poses = []
                for _ in range(4):
                    q = q0.copy()
                    q[active] = lo[active] + rng.random(len(active)) * (hi[active] - lo[active])
                    b.SetDOFValues(q.tolist())
                    poses.append(b.GetLinkTransformations())
                b.SetDOFValues(q0.tolist())
                for l in range(nl):
                    if not moves[l]:
                        assert all(_same(poses[0][l], p[l]) for p in poses[1:]), (f, b.name, l, active)
                rel = [[rm.t_inv(p[l]) for l in range(nl)] for p in poses]
                for k, (a, c) in enumerate(pairs):
                    if inv[k]:
                        assert all(_same(rm.t_mul(rel[0][a], poses[0][c]), rm.t_mul(rel[i][a], poses[i][c])) for i in range(1, 4)), (f, b.name, a, c, active)
            bodies += 1
    assert bodies >= 50
