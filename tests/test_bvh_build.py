"""The device-side BVH builder (openrave_b200/csrc/cudacollision.cu: Builder, binned surface-area splits) through its
device-free diagnostic entry: structural invariants the traversal kernels rely on, conservativeness of the boxes, and the
surface-area cost against the plain midpoint split it replaced.  (Verdicts never depend on the tree's shape — the GPU parity
tests cover that; this pins the shape itself on the CPU.)"""
import ctypes
import os
import subprocess
import sys

import numpy as np

from conftest import load_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(verts, tris):
    from openrave_b200 import checker
    lib = ctypes.CDLL(checker.LIB_PATH)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.cc_debug_build_bvh.argtypes = [vp, ci, vp, ci, vp, ci, ctypes.POINTER(ci)]
    v = np.ascontiguousarray(verts, dtype=np.float64)
    t = np.ascontiguousarray(tris, dtype=np.int32)
    cnt = ci(0)
    assert lib.cc_debug_build_bvh(v.ctypes.data, len(v), t.ctypes.data, len(t), None, 0, ctypes.byref(cnt)) == 0
    nodes = np.zeros((cnt.value, 8), dtype=np.float32)
    assert lib.cc_debug_build_bvh(v.ctypes.data, len(v), t.ctypes.data, len(t), nodes.ctypes.data, cnt.value, ctypes.byref(cnt)) == 0
    return nodes, nodes[:, 3].copy().view(np.int32)


def check_tree(verts, tris, nodes, left):
    nt = len(tris)
    assert len(nodes) == 2 * nt - 1
    c, e = nodes[:, 0:3].astype(np.float64), nodes[:, 4:7].astype(np.float64)
    seen = np.zeros(nt, dtype=np.int64)
    area = 0.0
    stack = [0]
    visited = 0
    while stack:
        i = stack.pop()
        visited += 1
        if left[i] < 0:
            t = ~left[i]
            seen[t] += 1
            p = verts[tris[t]]
            assert (p >= c[i] - e[i] - 1e-12).all() and (p <= c[i] + e[i] + 1e-12).all()      # the leaf box holds its triangle
            assert (e[i] <= 0.5 * (p.max(0) - p.min(0)) + 3e-5).all()                          # and is tight up to the inflation
            continue
        l = left[i]
        assert 0 < l and l + 1 < len(nodes) and l % 2 == 1                                     # sibling pairs are adjacent, 64-byte aligned
        for k in (l, l + 1):
            assert (c[k] - e[k] >= c[i] - e[i] - 2.5e-5).all() and (c[k] + e[k] <= c[i] + e[i] + 2.5e-5).all()   # children inside the parent
            stack.append(k)
        area += sum(2 * (e[k][0] * e[k][1] + e[k][1] * e[k][2] + e[k][2] * e[k][0]) * 4 for k in (l, l + 1))
    assert visited == len(nodes) and (seen == 1).all()                                         # every triangle in exactly one leaf
    assert np.allclose(nodes[:, 7], (e[:, 0] * e[:, 1] * e[:, 2]).astype(np.float32), rtol=1e-5)   # the descent's size key = volume
    return area


def test_invariants_on_real_links_and_degenerate_meshes():
    env = load_scene("lab1")
    links = [l for b in env.GetBodies() for l in b.links if len(l.idx) > 0]
    links.sort(key=lambda l: -len(l.idx))
    for l in links[:3] + links[-3:]:
        v, t = np.asarray(l.verts, dtype=np.float64), np.asarray(l.idx, dtype=np.int32)
        nodes, left = build(v, t)
        check_tree(v, t, nodes, left)
    # degenerate inputs: one triangle, two triangles, many identical triangles (all centroids equal: the median fallback)
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    for t in ([[0, 1, 2]], [[0, 1, 2], [0, 1, 3]], [[0, 1, 2]] * 37, [[0, 1, 2], [1, 2, 3], [0, 2, 3], [0, 1, 3]] * 5):
        t = np.array(t, dtype=np.int32)
        nodes, left = build(v, t)
        check_tree(v, t, nodes, left)
    # collinear centroids along a diagonal, wildly different triangle sizes
    rng = np.random.default_rng(1)
    n = 500
    base = np.linspace(0, 1, n)[:, None] * np.array([[1.0, 1.0, 1.0]])
    size = 10.0 ** rng.uniform(-4, -1, n)
    v = (base[:, None, :] + size[:, None, None] * rng.normal(size=(n, 3, 3))).reshape(-1, 3)
    t = np.arange(3 * n, dtype=np.int32).reshape(n, 3)
    nodes, left = build(v, t)
    check_tree(v, t, nodes, left)


def test_surface_area_cost_beats_the_midpoint_split():
    """CC_TUNE_BVH=mid selects the old rule (read once per process, hence the subprocess): on the scene's larger meshes the binned
    surface-area build must have the lower total child surface area — the quantity the split minimises level by level."""
    code = ("import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from conftest import load_scene\nfrom test_bvh_build import build, check_tree\n"
            "env = load_scene('lab1')\n"
            "links = sorted([l for b in env.GetBodies() for l in b.links if len(l.idx) > 500], key=lambda l: -len(l.idx))[:6]\n"
            "tot = 0.0\n"
            "for l in links:\n"
            "    v, t = np.asarray(l.verts, dtype=np.float64), np.asarray(l.idx, dtype=np.int32)\n"
            "    nodes, left = build(v, t)\n    tot += check_tree(v, t, nodes, left)\n"
            "print('AREA', tot)\n") % (ROOT, os.path.join(ROOT, "tests"))
    areas = {}
    for mode in ("sah", "mid"):
        env = dict(os.environ, CC_TUNE_BVH=mode)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
        assert "AREA" in out.stdout, out.stderr[-2000:]
        areas[mode] = float(out.stdout.split("AREA")[1])
    assert areas["sah"] < 0.95 * areas["mid"], areas
