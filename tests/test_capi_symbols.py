"""The C-ABI library loads and exports every symbol include/cudacollision.h declares (not gpu, no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cudacollision.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_header():
    from openrave_b200 import checker
    if not os.path.isfile(checker.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(checker.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    from openrave_b200 import planning
    assert sorted(checker.EXPORTS + planning.PLANNER_EXPORTS) == syms


def test_no_device_fails_loudly():
    """Without a CUDA device the product must refuse to work (no CPU path)."""
    import subprocess
    import sys
    code = ("import os,sys; os.environ['CUDA_VISIBLE_DEVICES']=''; sys.path.insert(0,%r);"
            "from openrave_b200.checker import CudaCollisionChecker, CudaCollisionError\n"
            "try:\n    CudaCollisionChecker(None,0); print('CREATED')\n"
            "except CudaCollisionError as e:\n    print('REFUSED', e)\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120).stdout
    assert "REFUSED" in out and "no CUDA device" in out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "openrave_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_header_is_plain_c():
    """The boundary is a C ABI: the header must compile as C99 (no C++ types, no torch types) and as C++11."""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "cudacollision.h")
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    for cmd in (["gcc", "-std=c99", "-Wall", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_planner_entries_validate_their_arguments():
    """The planner host needs no device: bad arguments come back as CC_ERR_INVALID / CC_ERR_STATE with a message, never a crash."""
    import numpy as np
    from openrave_b200 import planning as pl
    lib = pl._planner_lib()
    lib.cc_last_error.restype = ctypes.c_char_p
    n = 3
    lo, hi, res = np.full(n, -1.0), np.full(n, 1.0), np.full(n, 0.05)
    sp = pl.cc_cspace(n, 0, lo.ctypes.data, hi.ctypes.data, res.ctypes.data, None, None)
    h = ctypes.c_void_p()
    assert lib.cc_planner_create(None, ctypes.byref(h)) != 0 and b"bad configuration space" in lib.cc_last_error()
    assert lib.cc_planner_create(ctypes.byref(sp), ctypes.byref(h)) == 0
    try:
        prm = pl.cc_birrt_params(0.04, 0.01, 0.2, 100, 10, 0, 100, -8, 0)
        q = np.zeros((1, n))
        res1 = (pl.cc_plan_result * 1)()
        seeds = np.zeros(1, dtype=np.uint32)
        # no checker bound yet
        rc = lib.cc_birrt_plan_batch(h, ctypes.byref(prm), 1, q.ctypes.data, q.ctypes.data, seeds.ctypes.data, ctypes.addressof(res1))
        assert rc != 0 and b"no checker bound" in lib.cc_last_error()
        er = (pl.cc_edge_result * 1)()
        rc = lib.cc_planner_check_edges(h, 1, q.ctypes.data, q.ctypes.data, None, None, None, 3, 0xffff, 0.1, ctypes.addressof(er), None)
        assert rc != 0 and b"no checker bound" in lib.cc_last_error()
        # the samplers work without a checker: a straight edge of 0.5 along DOF 0 at resolution 0.05 has 10 steps
        cnt = ctypes.c_int(0)
        q1 = np.array([[0.5, 0.0, 0.0]])
        assert lib.cc_planner_edge_samples(h, q.ctypes.data, q1.ctypes.data, pl.IT_Closed, None, 0, ctypes.byref(cnt)) == 0
        assert cnt.value == 11                                   # q1, q0, steps 1..9
        assert lib.cc_planner_edge_samples(h, q.ctypes.data, q1.ctypes.data, 7, None, 0, ctypes.byref(cnt)) != 0
        v = np.zeros((1, n))
        assert lib.cc_planner_ramp_samples(h, q.ctypes.data, q1.ctypes.data, v.ctypes.data, v.ctypes.data, 0.0, pl.IT_Closed,
                                           None, None, None, 0, ctypes.byref(cnt)) != 0      # timeelapsed must be > 0
        # a checker that always answers "free"; step_length must be positive
        CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32),
                              ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.POINTER(ctypes.c_int32),
                              ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32))

        def free(user, qq, nconfigs, ndof, gb, ngroups, sb, nseg, ffs, fv, ffc):
            for g in range(ngroups):
                ffs[g], fv[g], ffc[g] = -1, 0, -1
            return 0
        cb = CB(free)
        assert lib.cc_planner_set_checker(h, ctypes.cast(cb, ctypes.c_void_p), None) == 0
        bad = pl.cc_birrt_params(0.0, 0.01, 0.2, 100, 10, 0, 100, -8, 0)
        rc = lib.cc_birrt_plan_batch(h, ctypes.byref(bad), 1, q.ctypes.data, q1.ctypes.data, seeds.ctypes.data, ctypes.addressof(res1))
        assert rc != 0 and b"step_length" in lib.cc_last_error()
        # in free space the plan is the straight line, found in the first iteration
        rc = lib.cc_birrt_plan_batch(h, ctypes.byref(prm), 1, q.ctypes.data, q1.ctypes.data, seeds.ctypes.data, ctypes.addressof(res1))
        assert rc == 0 and res1[0].status == pl.PLAN_SOLVED and res1[0].iterations == 1
        assert abs(res1[0].length_raw - 0.5) < 1e-9
        rc = lib.cc_planner_check_edges(h, 1, q.ctypes.data, q1.ctypes.data, None, None, None, pl.IT_Closed, 0xffff, 0.1, ctypes.addressof(er), None)
        assert rc == 0 and er[0].return_code == 0 and er[0].nsamples == 11 and er[0].first_invalid == -1
        rc = lib.cc_planner_check_edges(h, 1, q.ctypes.data, q1.ctypes.data, None, None, None, pl.IT_Closed, 1, 0.1, ctypes.addressof(er), None)
        assert rc != 0                                           # env and self are checked together
    finally:
        lib.cc_planner_destroy(h)
