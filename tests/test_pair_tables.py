"""Certified free-space tables (csrc/cudacollision.cu build_pair_tables): skipping a pair inside a certified cell must never
change a verdict or a reported pair.  GPU only: tables on vs tables off on 200k configurations (bitwise), and tables on vs
the oracle."""
import numpy as np
import pytest

import oracle
from conftest import ORACLE_KIND, load_scene, sample_configs


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
