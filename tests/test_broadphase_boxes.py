"""The broadphase kernel's box test (cc_broad_kernel: 15-bit fixed-point {min, max} triples, four packed subtractions) on the host.

The reference has no broadphase — `CheckCollisionP` visits every link pair (plugins/pqprave/collisionPQP.h:432-498) — so the only
thing a broadphase may do is drop a pair whose world AABBs are disjoint.  `cc_debug_broad_overlap` compiles the kernel's own
`quantise_box` / `qbox_overlap` for the host; these tests need no device.
"""
import ctypes

import numpy as np

from openrave_b200.checker import load_library


def _run(a, b):
    lib = load_library()
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.zeros(len(a), dtype=np.uint8)
    rc = lib.cc_debug_broad_overlap(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), len(a), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return (out & 1) != 0, (out & 2) != 0


def _boxes(rng, n, spread, emax):
    c = (rng.random((n, 3)) - 0.5) * 2 * spread
    e = rng.random((n, 3)) * emax
    out = np.zeros((n, 8), dtype=np.float32)
    out[:, 0:3] = c
    out[:, 4:7] = e
    return out


def test_an_overlap_of_the_fp32_boxes_is_never_lost():
    rng = np.random.default_rng(11)
    lost = 0
    for spread, emax in ((1.5, 0.3), (3.0, 0.05), (3.0, 1.0), (18.0, 2.0), (40.0, 5.0), (0.2, 0.001)):
        a, b = _boxes(rng, 300000, spread, emax), _boxes(rng, 300000, spread, emax)
        # b around a: per-axis offsets up to 1.3 x the sum of the half extents, so that about half of the pairs overlap and many
        # of the others miss by a hair
        b[:, 0:3] = a[:, 0:3] + (rng.random((300000, 3)).astype(np.float32) - 0.5) * 2.6 * (a[:, 4:7] + b[:, 4:7])
        exact, quant = _run(a, b)
        assert 0.2 < exact.mean() < 0.7
        lost += int((exact & ~quant).sum())
    assert lost == 0


def test_touching_boxes_and_the_edges_of_the_quantisation_cube():
    rng = np.random.default_rng(12)
    n = 200000
    a = _boxes(rng, n, 4.0, 0.5)
    b = a.copy()
    axis = rng.integers(0, 3, n)
    sign = rng.choice([-1.0, 1.0], n).astype(np.float32)
    eb = (rng.random((n, 3)) * 0.5).astype(np.float32)
    b[:, 4:7] = eb
    # b touches a exactly on one face (the FP32 sum may round either way: bit 0 says which), overlapping on the other axes
    idx = np.arange(n)
    b[idx, axis] = a[idx, axis] + sign * (a[idx, 4 + axis] + eb[idx, axis])
    exact, quant = _run(a, b)
    assert exact.mean() > 0.3 and not (exact & ~quant).any()
    # boxes far outside the 32 m cube saturate: still conservative, and a box outside never overlaps one well inside
    far = _boxes(rng, n, 0.5, 0.2)
    far[:, 0] += 40.0
    near = _boxes(rng, n, 2.0, 0.5)
    exact, quant = _run(far, near)
    assert not exact.any() and not quant.any()
    far2 = far.copy()
    far2[:, 0:3] += (rng.random((n, 3)).astype(np.float32) - 0.5) * 0.3
    exact, quant = _run(far, far2)
    assert exact.sum() > 1000 and not (exact & ~quant).any()
    # boxes spanning the whole cube, zero-size boxes, NaN-free degenerate inputs
    big = np.zeros((4, 8), dtype=np.float32); big[:, 4:7] = 100.0
    pts = np.zeros((4, 8), dtype=np.float32); pts[:, 0:3] = [[0, 0, 0], [15.99, 0, 0], [-16.5, 3, 3], [0, 0, 31]]
    exact, quant = _run(big, pts)
    assert exact.all() and quant.all()


def test_the_quantised_test_prunes_almost_as_much_as_the_exact_one():
    """0.98 mm steps: the extra candidates are pairs whose boxes are within ~2 mm of touching (a few per cent of the overlapping
    pairs for boxes of link size, each costing one root-box test in the traversal)."""
    rng = np.random.default_rng(13)
    a, b = _boxes(rng, 2000000, 1.5, 0.15), _boxes(rng, 2000000, 1.5, 0.15)
    exact, quant = _run(a, b)
    extra = (quant & ~exact).sum()
    assert exact.sum() > 1000 and extra <= 0.06 * exact.sum(), (int(extra), int(exact.sum()))
    # every extra candidate is a near miss: the largest per-axis gap is below two quantisation steps + slack
    gap = (np.abs(a[:, 0:3].astype(np.float64) - b[:, 0:3]) - (a[:, 4:7].astype(np.float64) + b[:, 4:7])).max(axis=1)
    assert (gap[quant & ~exact] < 2.5e-3).all()
