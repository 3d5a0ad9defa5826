#!/usr/bin/env python
"""Every scene, robot and kinbody file of the reference tree through the loaders, the flattening the checkers use, the oracle and
the host walk of the device's FK chain program: one line per file.  (Build container only: reads /root/reference.)

  python tools/scene_sweep.py > profiles/r02P_scene_sweep.txt
"""
import ctypes
import glob
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle                                                   # noqa: E402
from openrave_b200 import flatten                               # noqa: E402
from openrave_b200.checker import load_library                  # noqa: E402
from openrave_b200.environment import Environment               # noqa: E402
from openrave_b200.kinbody import external_mimic_values         # noqa: E402

REF = "/root/reference"


def fk_program_error(lib, b, rng):
    ext = []
    arr, nj = flatten.joint_table(b, externals=ext)
    lo, hi = (np.array(x) for x in b.GetDOFLimits())
    q = lo + (hi - lo) * rng.random((4, b.GetDOF()))
    rows = np.ascontiguousarray(np.hstack([q, external_mimic_values(b, q, ext)]) if ext else q)
    q0 = b.GetDOFValues()
    want = []
    for row in q:
        b.SetDOFValues(row.tolist())
        want.append(flatten.transforms8(b.GetLinkTransformations()))
    b.SetDOFValues(q0)
    want = np.array(want)
    nl = len(b.links)
    dof = np.array(q0, dtype=np.float64)
    base = flatten.transforms8([b.links[0].t])[0].copy()
    act = np.arange(b.GetDOF(), dtype=np.int32)
    worst = 0.0
    for walk in (0, 1):
        out = np.zeros((len(q), nl, 8))
        rc = lib.cc_debug_fk_program(ctypes.addressof(arr), nj, nl, b.GetDOF(), dof.ctypes.data_as(ctypes.c_void_p), base.ctypes.data_as(ctypes.c_void_p),
                                     act.ctypes.data_as(ctypes.c_void_p), len(act), None, walk, rows.ctypes.data_as(ctypes.c_void_p), len(q),
                                     out.ctypes.data_as(ctypes.c_void_p), None)
        if rc != 0:
            return "refused: " + lib.cc_last_error().decode(), len(ext)
        dq = np.minimum(np.abs(out[..., :4] - want[..., :4]).max(axis=-1), np.abs(out[..., :4] + want[..., :4]).max(axis=-1))
        worst = max(worst, dq.max(), np.abs(out[..., 4:7] - want[..., 4:7]).max())
    return "%.1e" % worst, len(ext)


def main():
    lib = load_library()
    rng = np.random.default_rng(1)
    files = []
    for pat in ("src/data/*", "src/robots/*", "src/collada_robots/*.zae", "test/testdata/*"):
        files += [f for f in sorted(glob.glob(os.path.join(REF, pat))) if os.path.isfile(f) and f.endswith((".xml", ".zae", ".dae")) and ".sensor." not in f]
    loaded = 0
    print("%-46s %-8s %s" % ("file", "status", "bodies: name links/dof/passive/triangles  [non-adjacent pairs, FK program vs SetDOFValues, external mimic joints]"))
    for f in files:
        rel = os.path.relpath(f, REF)
        env = Environment()
        try:
            env.Load(f)
        except Exception as e:
            print("%-46s %-8s %s: %s" % (rel, "REFUSED", type(e).__name__, str(e)[:110]))
            continue
        loaded += 1
        orc = oracle.OracleChecker(env, "port")
        env.SetCollisionChecker(orc)
        parts = []
        for b in env.GetBodies():
            s = "%s %d/%d/%d/%d" % (b.name, len(b.links), b.GetDOF(), len(b.passive_joints), sum(len(l.idx) for l in b.links))
            if b.GetDOF() > 0 and len(b.links) > 1:
                err, next_ = fk_program_error(lib, b, rng)
                s += " [%d, %s, %d]" % (len(b.GetNonAdjacentLinks(0)), err, next_)
            parts.append(s)
        orc.DestroyEnvironment()
        print("%-46s %-8s %s" % (rel, "ok", "; ".join(parts)[:400]))
    print("\n%d of %d files load" % (loaded, len(files)))


if __name__ == "__main__":
    main()
