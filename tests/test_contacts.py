"""CO_Contacts: plugins/pqprave/collisionPQP.h:578-603 runs PQP_ALL_CONTACTS and turns every intersecting
triangle pair into a CONTACT through OpenRAVE::geometry::TriTriCollision (include/openrave/geometry.h:1817-1912).

CPU: the product's contact routine (a host function of the library, no device needed) against the UNMODIFIED reference
routine compiled into oracle/_ref, on random and near-degenerate triangle pairs.
GPU: the intersecting triangle pairs the kernel enumerates == the contact list of the oracle's PQP_Collide(ALL_CONTACTS) for
every colliding link pair; positions / normals == the reference routine on those triangles."""
import ctypes

import numpy as np
import pytest

import oracle
from conftest import have_ref_oracle, load_scene, sample_configs
from openrave_b200 import ravemath as rm


def _product_contact(lib, u, v):
    u = np.ascontiguousarray(u, dtype=np.float64); v = np.ascontiguousarray(v, dtype=np.float64)
    pos, norm = np.zeros(3), np.zeros(3)
    rc = lib.cc_tri_tri_contact(u.ctypes.data, v.ctypes.data, pos.ctypes.data, norm.ctypes.data)
    return bool(rc), pos, norm


@pytest.mark.skipif(not have_ref_oracle(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_contact_routine_equals_reference_tritricollision():
    from openrave_b200.checker import load_library
    from openrave_b200.environment import Environment
    lib = load_library()
    ref = oracle.OracleChecker(Environment(), "reference")
    rng = np.random.default_rng(3)
    nhit = 0
    for k in range(6000):
        u = rng.normal(size=(3, 3))
        v = rng.normal(size=(3, 3)) * (0.3 if k % 3 else 1.0) + (u.mean(axis=0) if k % 2 else 0.0)
        if k % 10 == 0:                      # a shared vertex, an edge in the other's plane
            v[0] = u[0]
        if k % 17 == 0:
            v[:, 2] = u[:, 2].mean()
        a = _product_contact(lib, u, v)
        b = ref.tri_tri_collision(u, v)
        assert a[0] == b[0]
        if a[0]:
            nhit += 1
            assert np.abs(a[1] - b[1]).max() < 1e-12 and np.abs(a[2] - b[2]).max() < 1e-12
    assert 500 < nhit < 5500


@pytest.mark.gpu
def test_gpu_contact_lists_equal_oracle():
    from openrave_b200 import flatten
    from openrave_b200.checker import CO_Contacts, CudaCollisionChecker
    from openrave_b200.environment import CollisionReport
    env = load_scene("lab1")
    robot = env.GetRobots()[0]
    orc = oracle.OracleChecker(env, "port")
    env.SetCollisionChecker(orc)
    env2 = load_scene("lab1")
    robot2 = env2.GetRobots()[0]
    cc = CudaCollisionChecker(env2, 0)
    env2.SetCollisionChecker(cc)
    assert cc.SetCollisionOptions(CO_Contacts)
    ref = oracle.OracleChecker(load_scene("lab1"), "reference") if have_ref_oracle() else None
    q = sample_configs(robot, 400, seed=81)
    v = orc.check_batch(robot, q)["verdict"]
    checked = npairs = 0
    for row in q[(v & 1) != 0][:12]:
        robot.SetDOFValues(row.tolist())
        robot2.SetDOFValues(row.tolist())
        bq = cc.bind(robot2)
        T = flatten.transforms8(robot2.GetQueryLinkTransformations())
        pidx, tris, has, pos, norm, total = bq.contacts(T, flags=1)
        assert total == len(pidx) > 0
        by_pair = {}
        for p, (ta, tb) in zip(pidx, tris):
            by_pair.setdefault(int(p), set()).add((int(ta), int(tb)))
        bodies = orc._bodies_by_id()
        for p, got in by_pair.items():
            l1, ob, l2, _ = bq.pairs[p]
            n, _, ids = orc.collide_pair(robot.links[l1], bodies[ob].links[l2])
            want = set((int(a), int(b)) for a, b in ids)
            # identical sets up to triangle pairs that only touch within the FP32 error of the device test
            assert len(got ^ want) <= max(1, len(want) // 200), (len(got), len(want), len(got ^ want))
            npairs += 1
        if ref is not None:
            for i in range(0, len(pidx), max(1, len(pidx) // 40)):
                l1, ob, l2, _ = bq.pairs[pidx[i]]
                la, lb = robot.links[l1], bodies[ob].links[l2]
                u = np.array([rm.t_apply(la.t, la.verts[k]) for k in la.idx[tris[i, 0]]])
                w = np.array([rm.t_apply(lb.t, lb.verts[k]) for k in lb.idx[tris[i, 1]]])
                hit, rp, rn = ref.tri_tri_collision(u, w)
                assert hit == bool(has[i])
                if hit:
                    assert np.abs(rp - pos[i]).max() < 1e-9 and np.abs(rn - norm[i]).max() < 1e-9
                    checked += 1
        # the report path
        rep = CollisionReport()
        assert env2.CheckCollision(robot2, report=rep)
        assert sum(len(c) for c in rep.contacts_by_pair.values()) == int(has.sum())
        assert set(rep.contacts_by_pair) <= set(rep.vLinkColliding)
    assert npairs >= 12 and (ref is None or checked > 50)
