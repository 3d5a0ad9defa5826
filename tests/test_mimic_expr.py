"""The library's mimic-equation evaluator (openrave_b200/csrc/cc_mimic_expr.h) against the reference's own parser.

OpenRAVE evaluates mimic_pos strings with the vendored fparser 4.5.2 (kinbodyjoint.cpp:748-765, 2243-2249; kinbody.cpp:2308-2320).
`oracle/_ref/libfparser_ref.so` is that parser, built unmodified from where it lies in the reference tree (oracle/Makefile), with the
comparison epsilon OpenRAVE gives it.  Every expression below goes through both; values agree to rounding of the elementary
functions, evaluation errors carry the same code, and what fparser refuses to parse the library refuses too.  The second half pins
`cc_debug_expand_rows` (= `cc_query_expand_rows` on explicit arguments) against the harness' SetDOFValues restatement.
"""
import ctypes
import os

import numpy as np
import pytest

import oracle
from conftest import load_scene
from openrave_b200 import flatten
from openrave_b200.checker import load_library
from openrave_b200.kinbody import MimicExpr, external_mimic_values

FPREF = os.path.join(os.path.dirname(oracle.__file__), "_ref", "libfparser_ref.so")
needs_fparser = pytest.mark.skipif(not os.path.isfile(FPREF), reason="oracle/_ref/libfparser_ref.so not built (needs /root/reference)")


def _names(names):
    arr = (ctypes.c_char_p * max(len(names), 1))()
    for k, n in enumerate(names):
        arr[k] = n.encode()
    return arr


def ours(expr, names, values):
    lib = load_library()
    v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, max(len(names), 1))
    out = np.zeros(len(v))
    err = np.zeros(len(v), dtype=np.int32)
    rc = lib.cc_debug_mimic_eval(expr.encode(), _names(names), len(names), v.ctypes.data_as(ctypes.c_void_p), len(v),
                                 out.ctypes.data_as(ctypes.c_void_p), err.ctypes.data_as(ctypes.c_void_p))
    if rc != 0:
        return None, lib.cc_last_error().decode()
    return out, err


_fp = None


def theirs(expr, names, values):
    global _fp
    if _fp is None:
        _fp = ctypes.CDLL(FPREF)
        _fp.fpref_parse.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        _fp.fpref_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _fp.fpref_free.argtypes = [ctypes.c_void_p]
    h = ctypes.c_void_p()
    if _fp.fpref_parse(expr.encode(), ",".join(names).encode(), ctypes.byref(h)) != -1:
        return None, None
    v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1, max(len(names), 1))
    out = np.zeros(len(v))
    err = np.zeros(len(v), dtype=np.int32)
    _fp.fpref_eval(h, v.ctypes.data_as(ctypes.c_void_p), len(names), len(v), out.ctypes.data_as(ctypes.c_void_p), err.ctypes.data_as(ctypes.c_void_p))
    _fp.fpref_free(h)
    return out, err


VALID = [
    "x/3+0.8727", "-x", "0+x", "x*y", "x-y-1", "x/y/2", "2^3^2", "-x^2", "2^-x", "x^0.5", "x^2^0.5", "2*-x", "2--x", "--x", "-(-x)",
    "x%3", "-7%3", "7%-3", "x%y", "1e-3*x", ".5+x", "5.", "0x10+x", "1.5e+2*x", "(x+y)*(x-y)", "((x))", " x  +  1 ",
    "x=2", "x!=2", "x<2", "x<=2", "x>2", "x>=2", "x=y", "x<y", "x<2|x>3", "x>1&x<3", "x&1", "x|0", "!x", "!!x", "!x=0", "-x<0",
    "1+2<4", "x+1e-14=x", "x+1e-13=x", "x<x+1e-14", "x<=x-1e-14", "x<=x-1e-13", "0.49|0", "0.5|0", "!0.49", "!0.5", "-0.5&1",
    "if(x<1,10,20)", "if(x>1,10,20)", "if(x,y,-y)", "if(x<0,1/0,x)", "if(x>0,x,sqrt(-1))", "if(if(x,0,1),2,3)", "if(x>1,if(y>1,1,2),3)",
    "1+if(x>1,10,20)*2",
    "int(2.5)", "int(-2.5)", "int(3.49)", "int(x*1.3)", "trunc(-2.7)", "ceil(2.1)", "floor(-2.1)", "abs(-x)", "min(x,1)+max(x,3)",
    "min(x,y)", "max(x,y)", "sin(x)", "cos(x)", "tan(x)", "asin(x/4)", "acos(x/4)", "atan(x)", "atan2(1,x)", "atan2(y,x)", "sinh(x)",
    "cosh(x)", "tanh(x)", "asinh(x)", "acosh(1+x*x)", "atanh(x/4)", "exp(x)", "exp2(x)", "log(1+x*x)", "log10(1+x*x)", "log2(1+x*x)",
    "sqrt(x*x)", "cbrt(x)", "hypot(x,y)", "pow(x,2)", "pow(2,x)", "cot(x+0.3)", "sec(x)", "csc(x+0.3)",
    "sin(x)*cos(y)+sqrt(x*x+y*y)/3", "1/(1+exp(-x))", "0.5*(x+abs(x))", "x*x/5+y/7+0.3", "atan2(sin(x),cos(x))",
    "x^1.5", "x^-0.5", "x^y", "x^-1", "x^-2", "x^-3", "pow(x,0.5)", "pow(x,y)", "pow(x,-1)", "x^(1/3)", "x^(1/2)", "x^0.25", "x^2.5",
    "x^3", "x^0.5^2", "sqrt(x)^2", "x^(1-0.5)", "x^-1.5", "1/x^2", "x^0", "0^x", "y^x", "2^x^y", "if(x>0,2,0.5)+3", "x^if(y>0,2,0.5)",
    "if(x>0,1,2)*if(y>0,3,4)", "2+3*4-1", "(1+2)^(1/2)", "exp(x*300)",
    "1/0", "1/x", "x%0", "sqrt(-1)", "sqrt(x)", "log(0)", "log(x)", "log10(-1)", "log2(x-5)", "asin(2)", "acos(x)", "acosh(0)",
    "atanh(1)", "atanh(x)", "cot(0)", "csc(0)",
]

INVALID = [
    "", " ", "1+", "(x", "x)", "2x", "2(x)", "(x)(x)", "x e", "sin x", "sin(x)cos(x)", "3 4", "x,x", "+x", "-+x", "x<>2", "x&&1", "x||0",
    "pi", "z", "sin()", "sin(x,y)", "atan2(x)", "if(x,1)", "if(x)", "min(x)", "1.5e+2x", "x:=3", "x^", "*x", "x**2", "x!", "foo(x)",
]


@needs_fparser
def test_values_and_error_codes_match_fparser():
    rng = np.random.default_rng(7)
    vals = np.vstack([rng.uniform(-3.5, 3.5, (400, 2)), [[2.0, 1.0], [0.0, 0.0], [1.0, 1.0], [-1.0, 2.0], [0.5, -0.5], [3.0, 3.0], [1e-13, 0.0]]])
    for expr in VALID:
        a, ea = ours(expr, ["x", "y"], vals)
        b, eb = theirs(expr, ["x", "y"], vals)
        assert b is not None, expr                       # the list above is what fparser accepts
        assert a is not None, (expr, ea)
        # an evaluation error here <=> fparser reports one or returns NaN / inf (it rewrites powers at parse time, and whether a
        # bad power comes out as one of its codes or as a silent NaN depends on the rewrite); same code wherever both have one
        bad = (eb != 0) | ~np.isfinite(b)
        assert ((ea != 0) == bad).all(), (expr, vals[(ea != 0) != bad][:3], ea[(ea != 0) != bad][:3], eb[(ea != 0) != bad][:3])
        both = (ea != 0) & (eb != 0) & (ea != 5)
        assert (ea[both] == eb[both]).all(), (expr, ea[both][:3], eb[both][:3])
        ok = ~bad
        assert (a[~ok] == 0).all()
        # elementary functions may differ in the last places (fparser computes log2 as log * 1/ln 2, pow through exp/log for
        # some arguments); the mimic value feeds a joint angle whose budget is 1e-6 rad
        scale = np.maximum(1.0, np.abs(b[ok]))
        assert (np.abs(a[ok] - b[ok]) <= 1e-12 * scale).all(), (expr, np.abs(a[ok] - b[ok]).max())


@needs_fparser
def test_what_fparser_rejects_is_rejected():
    for expr in INVALID:
        b, _ = theirs(expr, ["x", "y"], [[1.0, 2.0]])
        a, msg = ours(expr, ["x", "y"], [[1.0, 2.0]])
        if expr in ("x:=3",):
            assert b is None
        assert a is None, expr
        assert "position" in msg
        if b is not None:
            pytest.fail("fparser accepts %r" % expr)
    # accepted by fparser, refused here by design (documented): inline variables
    b, _ = theirs("t:=x+1;t*2", ["x", "y"], [[1.0, 2.0]])
    a, msg = ours("t:=x+1;t*2", ["x", "y"], [[1.0, 2.0]])
    assert b is not None and a is None and "not" in msg


def test_joint_names_are_matched_as_whole_words_longest_first():
    names = ["JF1", "JF1mimic", "r_gripper_l_finger_joint", "j-1", "a.b"]
    v = np.array([[2.0, 3.0, 5.0, 7.0, 11.0]])
    assert ours("JF1mimic/2 + JF1", names, v)[0][0] == 3.0 / 2 + 2.0
    assert ours("JF1+JF1mimic*JF1", names, v)[0][0] == 2.0 + 3.0 * 2.0
    assert ours("0+r_gripper_l_finger_joint", names, v)[0][0] == 5.0
    assert ours("j-1*2", names, v)[0][0] == 14.0           # the reference substitutes names before it parses: "j-1" is one joint
    assert ours("a.b+1", names, v)[0][0] == 12.0
    assert ours("JF1x", names, v)[0] is None               # not a joint, not a function
    assert ours("sin(JF1)", names, v)[0][0] == np.sin(2.0)


def test_python_twin_evaluator_agrees():
    """The harness' MimicExpr (arithmetic only; a negative base with a fractional exponent is NaN there, -(|x|^y) in fparser)."""
    rng = np.random.default_rng(9)
    vals = rng.uniform(-2, 2, (200, 2))
    for expr in ["x/3+0.8727", "x*x/5 + y/7 + 0.3", "-x^2", "2^-x", "(x+y)*(x-y)/3", "(x*x+1)^2^0.5", "--x"]:
        a, _ = ours(expr, ["x", "y"], vals)
        m = MimicExpr(expr)
        b = m.eval_many({"x": vals[:, 0], "y": vals[:, 1]}, len(vals))
        assert np.abs(a - b).max() < 1e-13, expr


def _nonlinear_world():
    env = load_scene("lab1")
    robot = env.GetRobots()[0]
    byname = {j.name: j for j in robot.all_joints()}
    byname["JF1mimic"].mimic = MimicExpr("JF1*JF1/5 + JF2/7 + 0.3")
    byname["JF2mimic"].mimic = MimicExpr("JF1mimic/2 + JF3/5 + 0.1")
    robot.compute_internal_information()
    return env, robot


def _expand(robot, arr, nj, ext, active, q, naffine=0):
    lib = load_library()
    js = robot.sorted_joints_all
    eqs = (ctypes.c_char_p * nj)()
    for k, j in enumerate(js):
        eqs[k] = j.mimic.text.encode() if any(j is e for e in ext) else None
    names = _names([j.name for j in js])
    act = np.ascontiguousarray(active, dtype=np.int32)
    dof = np.array(robot.GetDOFValues(), dtype=np.float64)
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.zeros((len(q), len(active) + len(ext) + naffine))
    rc = lib.cc_debug_expand_rows(ctypes.addressof(arr), nj, robot.GetDOF(), dof.ctypes.data_as(ctypes.c_void_p), eqs, names,
                                  act.ctypes.data_as(ctypes.c_void_p), len(act), naffine, q.ctypes.data_as(ctypes.c_void_p), len(q),
                                  out.ctypes.data_as(ctypes.c_void_p))
    return rc, out


def test_expanded_rows_are_what_the_twin_computes():
    """`cc_query_expand_rows` and `BoundQuery.rows` produce the same rows: the latter is what tests/test_external_mimic.py ran on
    the GPU against the harness FK and the oracle."""
    env, robot = _nonlinear_world()
    ext = []
    arr, nj = flatten.joint_table(robot, externals=ext)
    lo, hi = (np.array(x) for x in robot.GetDOFLimits())
    rng = np.random.default_rng(5)
    ndof = robot.GetDOF()
    q0 = np.array(robot.GetDOFValues())
    for active, naffine in ((list(range(ndof)), 0), ([robot.GetJoint("JF1").dofindex, 0, 1, 3], 0), ([7, 8], 3), ([], 0)):
        n = 300
        q = lo[active] + (hi[active] - lo[active]) * rng.random((n, len(active)))
        q[::5] = lo[active] + (hi[active] - lo[active]) * np.round(rng.random((len(q[::5]), len(active)))) * 1.2      # past the limits
        aff = rng.uniform(-1, 1, (n, naffine))
        rc, out = _expand(robot, arr, nj, ext, active, np.hstack([q, aff]), naffine)
        assert rc == 0
        full = np.tile(q0, (n, 1))
        full[:, active] = q
        for j in robot.joints:
            if not j.circular:
                full[:, j.dofindex] = np.clip(full[:, j.dofindex], j.lower, j.upper)
        want = external_mimic_values(robot, full, ext)
        na = len(active)
        assert (out[:, :na] == q).all() and (out[:, na + len(ext):] == aff).all()
        assert np.abs(out[:, na:na + len(ext)] - want).max() < 1e-14
        assert (np.abs(want[:, 0]) > 0).any()


def test_expand_rows_refuses_what_it_cannot_evaluate():
    env, robot = _nonlinear_world()
    ext = []
    arr, nj = flatten.joint_table(robot, externals=ext)
    lib = load_library()
    active = list(range(robot.GetDOF()))
    q = np.array([robot.GetDOFValues()])
    # an equation that divides by a DOF that is 0 in this row
    byname = {j.name: j for j in robot.all_joints()}
    byname["JF1mimic"].mimic.text = "1/JF1"
    q[0, robot.GetJoint("JF1").dofindex] = 0.0
    rc, _ = _expand(robot, arr, nj, ext, active, q)
    assert rc != 0 and "division by zero" in lib.cc_last_error().decode()
    # a large batch is split over threads: the first failing row in ROW order is the one named, and rows before it are filled
    big = np.tile(np.array(robot.GetDOFValues()), (40000, 1))
    big[:, robot.GetJoint("JF1").dofindex] = 0.5
    big[[31000, 9000, 25000], robot.GetJoint("JF1").dofindex] = 0.0
    rc, out = _expand(robot, arr, nj, ext, active, big)
    assert rc != 0 and "row 9000 " in lib.cc_last_error().decode()
    assert (out[:9000, len(active)] == 2.0).all()
    # an unknown function
    byname["JF1mimic"].mimic.text = "polyroots2(JF1, 1, 2)"
    rc, _ = _expand(robot, arr, nj, ext, active, q)
    assert rc != 0 and "position" in lib.cc_last_error().decode()
    # no equation at all for an external joint
    js = robot.sorted_joints_all
    eqs = (ctypes.c_char_p * nj)()
    names = _names([j.name for j in js])
    act = np.array(active, dtype=np.int32)
    dof = np.array(robot.GetDOFValues())
    out = np.zeros((1, len(active) + 2))
    rc = lib.cc_debug_expand_rows(ctypes.addressof(arr), nj, robot.GetDOF(), dof.ctypes.data_as(ctypes.c_void_p), eqs, names,
                                  act.ctypes.data_as(ctypes.c_void_p), len(act), 0, q.ctypes.data_as(ctypes.c_void_p), 1,
                                  out.ctypes.data_as(ctypes.c_void_p))
    assert rc != 0 and "no equation" in lib.cc_last_error().decode()


def test_a_robot_with_fparser_functions_in_its_equations_loads_and_moves():
    """The harness evaluates what Python's arithmetic cannot through the library's compiler: the loaders accept such robots, the
    FK restatement moves them, and the rows for the device come out the same either way."""
    env = load_scene("lab1")
    robot = env.GetRobots()[0]
    byname = {j.name: j for j in robot.all_joints()}
    byname["JF1mimic"].mimic = MimicExpr("0.3 + 0.5*sin(JF1)^2 + if(JF2 > 1, 0.2, 0.1*JF2)")
    byname["JF2mimic"].mimic = MimicExpr("max(JF1mimic/2, 0.25) + atan2(JF3, 1)/5")
    assert byname["JF1mimic"].mimic.affine() is None and byname["JF1mimic"].mimic.varnames == ["JF1", "JF2"]
    robot.compute_internal_information()
    ext = []
    arr, nj = flatten.joint_table(robot, externals=ext)
    assert [j.name for j in ext] == ["JF1mimic", "JF2mimic"]
    d = {n: robot.GetJoint(n).dofindex for n in ("JF1", "JF2", "JF3")}
    q = np.array(robot.GetDOFValues())
    q[d["JF1"]], q[d["JF2"]], q[d["JF3"]] = 0.9, 1.3, 0.4
    robot.SetDOFValues(q.tolist())
    e1 = 0.3 + 0.5 * np.sin(0.9) ** 2 + 0.2
    e2 = max(min(max(e1, byname["JF1mimic"].lower), byname["JF1mimic"].upper) / 2, 0.25) + np.arctan2(0.4, 1) / 5
    active = list(range(robot.GetDOF()))
    rc, out = _expand(robot, arr, nj, ext, active, q[None])
    assert rc == 0
    assert abs(out[0, len(active)] - e1) < 1e-15 and abs(out[0, len(active) + 1] - e2) < 1e-15
    want = external_mimic_values(robot, q[None], ext)
    assert np.abs(want - out[:, len(active):]).max() < 1e-15
    with pytest.raises(ValueError):
        MimicExpr("polyroots2(JF1, 1, 2)")
