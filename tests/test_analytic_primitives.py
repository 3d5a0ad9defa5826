"""Analytic primitives — the fclrave-style geometry mode (plugins/fclrave/fclspace.cpp:475-602 builds exact
fcl::Box / Sphere / Cylinder shapes where pqprave sees their tessellation).  FCL is external and un-vendored ("parity unpinned"), so
the pins are (1) closed forms on random poses of primitive pairs, through the C ABI, and (2) the tessellated (pqprave) answer on the
reference's scenes: the exact shapes contain their tessellation, so analytic contacts are a superset, and a configuration may differ
only if the tessellated separation is below the tessellation's sagitta."""
import ctypes
import math

import numpy as np
import pytest

from conftest import ORACLE_KIND, load_scene, sample_configs
from openrave_b200 import kinbody as kb
from openrave_b200 import ravemath as rm

pytestmark = pytest.mark.gpu


def _two_prim_scene(ga, gb):
    """A one-link 'robot' carrying primitive ga and a static one-link body carrying gb, both with their tessellated meshes."""
    from openrave_b200.checker import CudaCollisionChecker
    from openrave_b200.environment import Environment
    env = Environment()
    bodies = []
    for name, g, robot in (("mover", ga, True), ("obstacle", gb, False)):
        body = kb.Robot(env) if robot else kb.KinBody(env)
        body.name = name
        l = kb.Link(body, "base")
        l.index = 0
        geom = kb.Geometry(g["type"] if g["type"] != "capsule" else "cylinder")
        geom.data = tuple(g["data"]) + (0.0,) * (3 - len(g["data"]))
        geom.init_collision_mesh()
        l.geometries = [geom]
        l.verts, l.idx = geom.verts.copy(), geom.idx.copy()
        l.prims = [dict(type=g["type"], t=[[1.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0]], data=list(g["data"]))]
        body.links.append(l)
        env.Add(body)
        bodies.append(body)
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    cc.SetAnalyticPrimitives(True)
    return env, bodies[0], bodies[1], cc


def _random_poses(n, seed, spread):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = (rng.random((n, 3)) - 0.5) * 2 * spread
    T = np.zeros((n, 1, 8))
    T[:, 0, 0:4] = q
    T[:, 0, 4:7] = t
    return T


def _rot(q):
    return np.array(rm.matrix_from_quat(tuple(q))).reshape(3, 3)


def _closest_point_on_box(p, h):
    return np.clip(p, -h, h)


def _seg_seg_dist(p1, q1, p2, q2):
    """distance between segments [p1,q1] and [p2,q2] (Ericson, Real-Time Collision Detection 5.1.9)"""
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f = d1 @ d1, d2 @ d2, d2 @ r
    if a <= 1e-18 and e <= 1e-18:
        return np.linalg.norm(r)
    if a <= 1e-18:
        s, t = 0.0, np.clip(f / e, 0, 1)
    else:
        c = d1 @ r
        if e <= 1e-18:
            t, s = 0.0, np.clip(-c / a, 0, 1)
        else:
            b = d1 @ d2
            den = a * e - b * b
            s = np.clip((b * f - c * e) / den, 0, 1) if den > 1e-18 else 0.0
            t = (b * s + f) / e
            if t < 0:
                t, s = 0.0, np.clip(-c / a, 0, 1)
            elif t > 1:
                t, s = 1.0, np.clip((b - c) / a, 0, 1)
    return np.linalg.norm((p1 + d1 * s) - (p2 + d2 * t))


def _check(ga, gb, exact_distance, n=6000, spread=0.3, band=2e-5):
    env, mover, obstacle, cc = _two_prim_scene(ga, gb)
    T = _random_poses(n, 5, spread)
    bq = cc.bind(mover)
    assert bq.info.npairs_env == 1
    v, fe, _, _ = bq.check_poses(T)
    hit = (v & 1) != 0
    d = np.array([exact_distance(_rot(T[i, 0, 0:4]), T[i, 0, 4:7]) for i in range(n)])
    clear = np.abs(d) > band
    assert (hit[clear] == (d[clear] < 0)).all(), int((hit[clear] != (d[clear] < 0)).sum())
    assert 0.05 < hit.mean() < 0.95, hit.mean()
    assert (fe[hit] == 0).all() and (fe[~hit] == -1).all()
    return hit


def test_sphere_sphere_and_sphere_box_closed_forms():
    ra, rb = 0.12, 0.2
    _check(dict(type="sphere", data=[ra]), dict(type="sphere", data=[rb]), lambda R, t: np.linalg.norm(t) - ra - rb)
    h = np.array([0.25, 0.1, 0.18])
    # the sphere moves, the box stays: distance from the sphere centre to the box
    _check(dict(type="sphere", data=[ra]), dict(type="box", data=list(h)),
           lambda R, t: (np.linalg.norm(t - _closest_point_on_box(t, h)) - ra) if (np.abs(t) > h).any() else -1.0)
    # the box moves (rotated), the sphere stays at the origin: the same distance in the box's frame
    _check(dict(type="box", data=list(h)), dict(type="sphere", data=[ra]),
           lambda R, t: (np.linalg.norm((-R.T @ t) - _closest_point_on_box(-R.T @ t, h)) - ra) if (np.abs(R.T @ t) > h).any() else -1.0)


def test_capsule_capsule_and_sphere_cylinder_closed_forms():
    ra, ha, rb, hb = 0.06, 0.5, 0.09, 0.4                 # radius, height of the two capsules (axis z)
    za, zb = np.array([0, 0, ha / 2]), np.array([0, 0, hb / 2])
    _check(dict(type="capsule", data=[ra, ha]), dict(type="capsule", data=[rb, hb]),
           lambda R, t: _seg_seg_dist(R @ -za + t, R @ za + t, -zb, zb) - ra - rb)
    r, hc, rs = 0.15, 0.5, 0.08

    def sphere_cyl(R, t):                                 # sphere (moving) against the z cylinder at the origin
        rad, z = math.hypot(t[0], t[1]), abs(t[2])
        dr, dz = rad - r, z - hc / 2
        if dr <= 0 and dz <= 0:
            return -1.0
        return math.hypot(max(dr, 0.0), max(dz, 0.0)) - rs
    _check(dict(type="sphere", data=[rs]), dict(type="cylinder", data=[r, hc]), sphere_cyl)


def test_box_box_equals_the_tessellated_answer():
    """A box IS its tessellation: the analytic mode and the triangle mode must agree on box pairs outside the rounding band."""
    from openrave_b200.checker import BATCH_ENV
    ha, hb = [0.2, 0.05, 0.3], [0.1, 0.25, 0.15]
    env, mover, obstacle, cc = _two_prim_scene(dict(type="box", data=ha), dict(type="box", data=hb))
    T = _random_poses(8000, 7, 0.35)
    va = cc.bind(mover).check_poses(T)[0]
    cc.SetAnalyticPrimitives(False)
    bq = cc.bind(mover)
    vt = bq.check_poses(T)[0]
    d, _, _ = bq.distance(poses=T, flags=BATCH_ENV)
    # one box inside the other has no triangle contact: the analytic mode (solid shapes) reports it, the surface test cannot
    inside = (va != 0) & (vt == 0) & (d > 1e-3)
    differ = (va != vt) & ~inside
    assert (d[differ] < 2e-5).all(), (int(differ.sum()), float(d[differ].max()) if differ.any() else 0.0)
    assert 0.1 < (vt != 0).mean() < 0.9


@pytest.mark.parametrize("scene", ["wamtest1", "lab1"])
def test_scene_analytic_is_a_superset_within_the_sagitta(scene):
    from openrave_b200.checker import BATCH_ENV, BATCH_SELF, CudaCollisionChecker
    env = load_scene(scene)
    robot = env.GetRobots()[0]
    cc = CudaCollisionChecker(env, 0)
    env.SetCollisionChecker(cc)
    q = sample_configs(robot, 40000, seed=41)
    tess = cc.CheckCollisionBatch(robot, q)
    dist = cc.DistanceBatch(robot, q)["distance"]
    cc.SetAnalyticPrimitives(True)
    ana = cc.CheckCollisionBatch(robot, q)
    nprim = sum(1 for b in env.GetBodies() for l in b.links if getattr(l, "prims", None))
    assert nprim >= 8
    # the exact shapes contain their tessellation: every tessellated contact is an analytic contact ...
    lost = (tess["verdict"] & ~ana["verdict"]) != 0
    assert not lost.any(), int(lost.sum())
    # ... and a configuration gains a contact only if it was within the sagitta of the largest tessellated cylinder / sphere
    rmax = max([g["data"][0] for b in env.GetBodies() for l in b.links for g in (getattr(l, "prims", None) or []) if g["type"] != "box"] + [0.0])
    sagitta = rmax * (1.0 - math.cos(math.pi / 51.0)) + 2e-5
    gained = ana["verdict"] != tess["verdict"]
    assert (dist[gained] <= sagitta).all(), (int(gained.sum()), float(dist[gained].max()) if gained.any() else 0.0, sagitta)
    # the first colliding pair in reference order can only move EARLIER (an exact shape touching where its tessellation did not)
    for key in ("first_env_index", "first_self_index"):
        a, t = ana[key], tess[key]
        both = (a >= 0) & (t >= 0)
        assert ((t >= 0) <= (a >= 0)).all()
        assert (a[both] == t[both]).mean() > 0.98
        # ... except where the tessellated contact of the earlier pair is a graze: the FP32 GJK of the exact shapes and the FP32
        # triangle test may disagree inside the contact band of THAT pair (oracle: colliding with a penetration measure <= 1e-3 m)
        later = np.flatnonzero(both & (a > t))
        assert len(later) <= max(2, int(both.sum()) // 200), len(later)
        if len(later):
            import oracle
            env_o = load_scene(scene)
            orc = oracle.OracleChecker(env_o, ORACLE_KIND)
            orc._sync()          # registers the bodies (fills orc._ids)
            bq = tess["query"]
            for c in later[:20]:
                l1, ob, l2, _ = bq.pairs[int(t[c])]
                name = cc.body_from_id(int(ob)).name
                other = [b for b in env_o.GetBodies() if b.name == name][0]
                cls = orc.pair_band(env_o.GetRobots()[0], q[c:c + 1], [(int(l1), orc._ids[id(other)], int(l2))], band=1e-3)[0, 0]
                assert cls in (1, 2), "configuration %d: pair %d found by the triangle test, missed by the exact shapes, class %d" % (c, int(t[c]), cls)
    # complete masks agree where the verdicts do; single-state overloads run through the same pass
    m1 = cc.CheckCollisionBatch(robot, q[:2000], all_pairs=True)["mask"]
    cc.SetAnalyticPrimitives(False)
    m0 = cc.CheckCollisionBatch(robot, q[:2000], all_pairs=True)["mask"]
    assert ((m0 & ~m1) == 0).all()
    print(scene, "primitive links", nprim, "configs gaining a contact", int(gained.sum()), "sagitta", sagitta)
