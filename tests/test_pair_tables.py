"""Certified free-space tables (csrc/cudacollision.cu build_pair_tables): skipping a pair inside a certified cell must never
change a verdict or a reported pair.  GPU only: tables on vs tables off on 200k configurations (bitwise), and tables on vs
the oracle."""
import numpy as np
import pytest

import oracle
from conftest import ORACLE_KIND, load_scene, sample_configs  # noqa: F401


def _world(scene, manip, tables):
    from openrave_b200.checker import CudaCollisionChecker
    env = load_scene(scene)
    robot = env.GetRobots()[0]
    active = None
    if manip is not None:
        active = [m for m in robot.GetManipulators() if m.GetName() == manip][0].GetArmIndices()
        robot.SetActiveDOFs(active)
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    cc.SetPairTables(tables)
    return env, robot, cc, active


@pytest.mark.gpu
@pytest.mark.parametrize("scene,manip,expect_tables", [("pr2test1", "rightarm_torso", True), ("lab1", None, False),
                                                       ("hironxtable", None, False)])
def test_tables_do_not_change_results(scene, manip, expect_tables):
    env, robot, on, active = _world(scene, manip, True)
    env2, robot2, off, _ = _world(scene, manip, False)
    q = sample_configs(robot, 200000, seed=51, active=active)
    a = on.CheckCollisionBatch(robot, q, active=active)          # >= 16k configurations: the tables are built on this call
    ntab = a["query"].table_count()
    if expect_tables:
        assert ntab > 0
    b = off.CheckCollisionBatch(robot2, q, active=active)
    assert b["query"].table_count() == 0
    assert np.array_equal(a["verdict"], b["verdict"])
    assert np.array_equal(a["first_env_index"], b["first_env_index"])
    assert np.array_equal(a["first_self_index"], b["first_self_index"])
    # small batches afterwards take the split kernel with the permuted table index
    for n in (1, 37, 600, 5000):
        a2 = on.CheckCollisionBatch(robot, q[:n], active=active)
        assert np.array_equal(a2["verdict"], b["verdict"][:n]) and np.array_equal(a2["first_env_index"], b["first_env_index"][:n])
        assert np.array_equal(a2["first_self_index"], b["first_self_index"][:n])
    # adversarial inputs for the cell lookup: values exactly on / beyond the joint limits (SetDOFValues clamps them), circular
    # DOFs several turns away (not clamped: the lookup wraps them), values on cell boundaries
    idx = list(range(robot.GetDOF())) if active is None else list(active)
    lo = np.array([robot.joints[d].lower for d in idx]); hi = np.array([robot.joints[d].upper for d in idx])
    circ = np.array([bool(robot.joints[d].circular) for d in idx])
    rng = np.random.default_rng(5)
    qa = q[:60000].copy()
    pick = rng.random(qa.shape)
    qa = np.where(pick < 0.05, lo, qa)
    qa = np.where((pick >= 0.05) & (pick < 0.10), hi, qa)
    qa = np.where((pick >= 0.10) & (pick < 0.15) & ~circ, hi + 0.3, qa)
    qa = np.where((pick >= 0.15) & (pick < 0.20) & ~circ, lo - 0.3, qa)
    cells = lo + np.round((qa - lo) / (hi - lo) * 256) * (hi - lo) / 256            # on 256-cell boundaries
    qa = np.where((pick >= 0.20) & (pick < 0.30), cells, qa)
    turns = rng.integers(-2, 3, size=qa.shape) * (2 * np.pi)
    qa = np.where(circ, qa + turns, qa)
    a3 = on.CheckCollisionBatch(robot, qa, active=active)
    b3 = off.CheckCollisionBatch(robot2, qa, active=active)
    assert np.array_equal(a3["verdict"], b3["verdict"]) and np.array_equal(a3["first_env_index"], b3["first_env_index"])
    assert np.array_equal(a3["first_self_index"], b3["first_self_index"])
    # complete pair masks too
    m1 = on.CheckCollisionBatch(robot, q[:3000], active=active, all_pairs=True)["mask"]
    m2 = off.CheckCollisionBatch(robot2, q[:3000], active=active, all_pairs=True)["mask"]
    assert np.array_equal(m1, m2)
    # a conservative triangle margin set afterwards: certificates proven without it are dropped and rebuilt with it
    on.SetTriangleEpsilon(2e-3)
    off.SetTriangleEpsilon(2e-3)
    a4 = on.CheckCollisionBatch(robot, q[:40000], active=active)
    b4 = off.CheckCollisionBatch(robot2, q[:40000], active=active)
    assert np.array_equal(a4["verdict"], b4["verdict"]) and np.array_equal(a4["first_env_index"], b4["first_env_index"])
    assert np.array_equal(a4["first_self_index"], b4["first_self_index"])
    assert (a4["verdict"] != a["verdict"][:40000]).any() or scene != "pr2test1"      # the margin does change some verdicts
    a5 = on.CheckCollisionBatch(robot, q[:40000], active=active)                      # certificates back, with the larger margin
    assert np.array_equal(a5["verdict"], b4["verdict"])
    on.SetTriangleEpsilon(0.0)
    off.SetTriangleEpsilon(0.0)
    # and against the oracle
    env3 = load_scene(scene)
    robot3 = env3.GetRobots()[0]
    if active is not None:
        robot3.SetActiveDOFs(active)
    orc = oracle.OracleChecker(env3, ORACLE_KIND)
    env3.SetCollisionChecker(orc)
    ref = orc.check_batch(robot3, q[:2500], active=active, want_band=True)
    band = (ref["near"] != 0) | ((ref["verdict"] != 0) & (ref["penetration"] <= 1e-4))
    a6 = on.CheckCollisionBatch(robot, q[:2500], active=active)
    assert not ((a6["verdict"] != ref["verdict"]) & ~band).any()
    print(scene, "tables:", ntab)


def _link_rel_points(robot, q_full, a, b):
    """vertices of link b expressed in link a's frame (b = None: vertices of a in the world) at DOF values q_full"""
    from openrave_b200 import ravemath as rm
    robot.SetDOFValues(list(q_full))
    ta = robot.links[a].t
    if b is None:
        return np.array([rm.t_apply(ta, v) for v in robot.links[a].verts[::7]])
    rel = rm.t_mul(rm.t_inv(ta), robot.links[b].t)
    return np.array([rm.t_apply(rel, v) for v in robot.links[b].verts[::7]])


@pytest.mark.parametrize("scene,manip", [("pr2test1", "rightarm_torso"), ("lab1", None)])
def test_displacement_rates_bound_finite_differences(scene, manip):
    """Device-free pin of the certificates' Lipschitz constants (csrc/cc_certify.inl pair_rates, through cc_debug_pair_rates):
    for pairs that depend on one or two active DOFs, no vertex of one link moves faster in the other link's frame than K_d per
    unit of DOF d — checked by finite differences of the harness FK at random values of ALL the joints."""
    import ctypes
    from openrave_b200 import checker, flatten, ravemath as rm
    if not hasattr(rm, "t_apply"):
        pytest.skip("ravemath.t_apply missing")
    lib = checker.load_library()
    env = load_scene(scene)
    robot = env.GetRobots()[0]
    ndof = robot.GetDOF()
    active = list(range(ndof)) if manip is None else list([m for m in robot.GetManipulators() if m.GetName() == manip][0].GetArmIndices())
    cols = np.full(ndof, -1, dtype=np.int32)
    cols[active] = np.arange(len(active))
    arr, nj = flatten.joint_table(robot)
    nl = len(robot.links)
    radius = [float(np.sqrt((np.asarray(l.verts) ** 2).sum(axis=1).max())) if len(l.verts) else 0.0 for l in robot.links]
    lo, hi = (np.array(x) for x in robot.GetDOFLimits())
    q0 = np.array(robot.GetDOFValues())
    rng = np.random.default_rng(17)
    with_mesh = [l.index for l in robot.links if len(l.verts)]
    tested = 0
    for _ in range(4000):
        a = int(rng.choice(with_mesh))
        b = -1 if rng.random() < 0.3 else int(rng.choice(with_mesh))
        if a == b:
            continue
        dofs = (ctypes.c_int32 * 2)()
        K = (ctypes.c_double * 2)()
        n = ctypes.c_int(0)
        assert lib.cc_debug_pair_rates(ctypes.addressof(arr), nj, nl, cols.ctypes.data_as(ctypes.c_void_p), ndof, a, b, radius[a],
                                       radius[b] if b >= 0 else 0.0, dofs, K, ctypes.byref(n)) == 0
        if n.value == 0:
            continue
        for trial in range(3):
            q = q0.copy()
            q[active] = lo[active] + rng.random(len(active)) * (hi[active] - lo[active])
            base = _link_rel_points(robot, q, a, b if b >= 0 else None)
            for i in range(n.value):
                d = dofs[i]
                h = 1e-4 * (1 if q[d] + 1e-4 <= hi[d] else -1)
                q2 = q.copy()
                q2[d] += h
                moved = _link_rel_points(robot, q2, a, b if b >= 0 else None)
                speed = np.linalg.norm(moved - base, axis=1).max() / abs(h)
                assert speed <= K[i] * (1 + 1e-6) + 1e-9, (robot.links[a].name, b, d, speed, K[i])
        tested += 1
        if tested >= 60:
            break
    robot.SetDOFValues(list(q0))
    assert tested >= 20


@pytest.mark.gpu
def test_tightest_certified_cells_at_their_corners():
    """Adversarial inputs for the certificates: for the certified pairs with the least slack, the corners of their
    tightest cells — the points of a cell farthest from the centre the proof was made at — with every other active DOF random.
    The oracle's exact distance of that pair must still clear the contact band there, and tables on == tables off bitwise."""
    env, robot, on, active = _world("pr2test1", "rightarm_torso", True)
    env2, robot2, off, _ = _world("pr2test1", "rightarm_torso", False)
    lo = np.array([robot.joints[d].lower for d in active]); hi = np.array([robot.joints[d].upper for d in active])
    rng = np.random.default_rng(23)
    warm = lo + rng.random((20000, len(active))) * (hi - lo)
    a = on.CheckCollisionBatch(robot, warm, active=active)
    bq = a["query"]
    certs = bq.certificates()
    assert len(certs) > 100 and bq.table_count() > 0
    certs.sort(key=lambda c: c.min_slack)
    orc = oracle.OracleChecker(load_scene("pr2test1"), ORACLE_KIND)
    robot3 = orc.env.GetRobots()[0]
    robot3.SetActiveDOFs(active)
    orc.env.SetCollisionChecker(orc)
    rows, owner = [], []
    for ci, c in enumerate(certs[:24]):
        i0, i1 = (c.tightest_cell // c.n1, c.tightest_cell % c.n1) if c.col1 >= 0 else (c.tightest_cell, 0)
        for k0 in (0, 1):
            for k1 in ((0, 1) if c.col1 >= 0 else (0,)):
                for _ in range(6):
                    q = lo + rng.random(len(active)) * (hi - lo)
                    q[c.col0] = min(hi[c.col0], c.lo0 + (i0 + k0) * c.width0)
                    if c.col1 >= 0:
                        q[c.col1] = min(hi[c.col1], c.lo1 + (i1 + k1) * c.width1)
                    rows.append(q); owner.append(ci)
    rows = np.array(rows)
    assert all(c.min_slack > 0 for c in certs)
    for ci, c in enumerate(certs[:24]):
        sel = rows[np.array(owner) == ci]
        l1, ob, l2, cls = bq.pairs[c.pair]
        triple = (int(l1), orc._ids[id(robot3)] if cls == 1 else orc._ids[id([b for b in orc.env.GetBodies() if b.name == on.body_from_id(int(ob)).name][0])], int(l2))
        d, _ = orc.pair_queries(robot3, sel, [triple], active=active, cap=1.0)
        # proven: distance(centre) > delta + margin, so every point of the cell is farther than `margin` (3e-4 m) from contact
        assert d.min() > c.margin - 2e-5, (c.pair, d.min(), c.min_slack, c.delta)
    x = on.CheckCollisionBatch(robot, rows, active=active)
    y = off.CheckCollisionBatch(robot2, rows, active=active)
    assert np.array_equal(x["verdict"], y["verdict"]) and np.array_equal(x["first_env_index"], y["first_env_index"])
    assert np.array_equal(x["first_self_index"], y["first_self_index"])
