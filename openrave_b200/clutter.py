"""Synthetic clutter scene of BASELINE.json configs[4] ("C5"): the Barrett WAM of data/lab1.env.xml,
7 arm DOFs active, inside a 2 m cube of `nobstacles` sphere-like triangle meshes of ~`tris_per_obstacle` triangles each
(radius U[0.05, 0.15] m, centres uniform in the cube but outside a 0.3 m cylinder around the arm's base axis).

The meshes are geodesic spheres: every icosahedron face is cut into f x f triangles and the points are pushed onto the
sphere (20 f^2 triangles; f = 22 -> 9 680, the closest to the 10 k the workload names; plain 4-way icosphere subdivision
only offers 5 120 or 20 480).  Every obstacle is its own body with its own vertex array, radius and random orientation, so
the device holds `nobstacles` distinct BVHs (200 x 9 680 triangles = 93 MB of triangles + 124 MB of nodes: larger than the
126 MB L2, which is the point of this workload).

Everything is generated from a seed at run time (the GPU box has no /root/reference and a 200 MB fixture would not belong
in the repository); the robot comes from the committed lab1 fixture.
"""
import math
import os

import numpy as np

from . import ravemath as rm
from .kinbody import _ICO_X, _ICO_Y, _ICO_Z, Geometry, KinBody, Link


def geodesic_sphere(freq):
    """Unit geodesic sphere: (vertices float64 [nv,3], triangles int32 [20 freq^2, 3]); vertices are per-face (not welded
    across icosahedron edges), winding outward."""
    X, Y, Z = _ICO_X, _ICO_Y, _ICO_Z
    base = np.array([(Z, X, -Y), (X, Y, Z), (Y, Z, -X), (Y, Z, X), (X, -Y, Z), (Z, X, Y),
                     (-Y, Z, X), (Z, -X, -Y), (-X, Y, Z), (-Y, Z, -X), (-X, -Y, Z), (Z, -X, Y)], dtype=np.float64)
    ind = np.array([0, 1, 2, 1, 3, 4, 3, 5, 6, 2, 4, 7, 5, 6, 8, 2, 7, 9, 0, 5, 8, 7, 9, 10, 0, 1, 5, 7, 10, 11,
                    1, 3, 5, 6, 10, 11, 3, 6, 11, 9, 10, 8, 3, 4, 11, 6, 8, 10, 4, 7, 11, 1, 2, 4, 0, 8, 9, 0, 2, 9]).reshape(-1, 3)
    f = int(freq)
    # lattice points (i, j), i + j <= f, numbered row by row
    ij = [(i, j) for i in range(f + 1) for j in range(f + 1 - i)]
    num = {p: k for k, p in enumerate(ij)}
    tri_local = []
    for i in range(f):
        for j in range(f - i):
            tri_local.append((num[(i, j)], num[(i + 1, j)], num[(i, j + 1)]))
            if i + j + 1 < f:
                tri_local.append((num[(i + 1, j)], num[(i + 1, j + 1)], num[(i, j + 1)]))
    tri_local = np.array(tri_local, dtype=np.int32)
    ij = np.array(ij, dtype=np.float64) / f
    verts, tris = [], []
    for k, (a, b, c) in enumerate(ind):
        va, vb, vc = base[a], base[b], base[c]
        p = va[None, :] + ij[:, :1] * (vb - va)[None, :] + ij[:, 1:2] * (vc - va)[None, :]
        p /= np.linalg.norm(p, axis=1)[:, None]
        t = tri_local.copy()
        if np.dot(np.cross(vb - va, vc - va), va) < 0:
            t = t[:, ::-1]
        verts.append(p)
        tris.append(t + k * len(ij))
    return np.concatenate(verts), np.ascontiguousarray(np.concatenate(tris), dtype=np.int32)


def _random_unit_quat(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return tuple(float(x) for x in q)


def make_clutter_environment(lab1_fixture, nobstacles=200, freq=22, seed=5, cube=2.0, keepout=0.3):
    """Returns (env, robot).  The robot is lab1's BarrettWAM with its 7 arm DOFs active; no checker is set."""
    from .environment import Environment
    src = Environment()
    src.Load(lab1_fixture)
    robot = src.GetRobots()[0]
    env = Environment()
    # move the robot over: a fresh Environment with only the robot, then the obstacles (environment order = creation order)
    src.bodies.remove(robot)
    env.Add(robot)
    robot.SetActiveDOFs(robot.GetActiveManipulator().GetArmIndices())
    robot.GetLink("segway").Enable(False)      # "7-DOF serial arm (WAM arm links)": the mobile base is not part of the workload
    shoulder = np.array(robot.GetLink("wam1").GetTransform()[1], dtype=np.float64)
    rng = np.random.default_rng(seed)
    unit_v, unit_i = geodesic_sphere(freq)
    for k in range(nobstacles):
        while True:
            c = shoulder + (rng.random(3) - 0.5) * cube
            if math.hypot(c[0] - shoulder[0], c[1] - shoulder[1]) >= keepout:
                break
        r = 0.05 + 0.10 * rng.random()
        body = KinBody(env, "clutter%03d" % k)
        link = Link(body, "base")
        link.index = 0
        link.static = True
        g = Geometry("trimesh")
        g.verts = unit_v * r
        g.idx = unit_i
        link.geometries = [g]
        link.update_collision()
        link.t = (_random_unit_quat(rng), (float(c[0]), float(c[1]), float(c[2])))
        body.links.append(link)
        env.Add(body)
    return env, robot


def default_fixture():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenes", "lab1.npz")
