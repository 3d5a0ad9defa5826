"""Host-side Python mirror of the parts of OpenRAVE's KinBody the collision path depends on.

The reference computes all of this inside libopenrave (which cannot be built in this image, see
DESIGN.md); an OpenRAVE-side plugin would read the derived quantities straight off a live
`KinBody`.  For the standalone harness (tests, bench, fixtures) this module restates them so a
scene can be flattened into exactly what the C ABI (include/cudacollision.h) consumes:

  * joint frames        Joint._ComputeJointInternalInformation   src/libopenrave/kinbodyjoint.cpp:1343-1525
  * static-joint fold   Joint._ComputeInternalStaticInformation  kinbodyjoint.cpp:1527-1551
  * hierarchy / order   KinBody._ComputeInternalInformation      src/libopenrave/kinbody.cpp:4315-4860 (tree-structured bodies)
  * adjacency           kinbody.cpp:4940-5010, index helper :49-58
  * forward kinematics  KinBody.SetDOFValues                     kinbody.cpp:2121-2485
  * non-adjacent pairs  KinBody.GetNonAdjacentLinks              kinbody.cpp:5512-5600 (+ robot.cpp:2023-2084)
  * link collision mesh Link._Update                             src/libopenrave/kinbodylink.cpp:1202-1232
  * primitive meshes    GeometryInfo.InitCollisionMesh           src/libopenrave/kinbodygeometry.cpp:131-271, 513-575

All arithmetic is scalar IEEE double in the reference's association order (see ravemath.py).
Only tree-structured bodies (no closed kinematic loops) with 1-DOF hinge/slider joints are
supported; anything else raises.  This is load-time host logic, not the per-config hot path.
"""
import ast
import math

import numpy as np

from . import ravemath as rm

# dReal = double: g_fEpsilon = 1e-15 (include/openrave/openrave.h:130), the derived tolerances are RavePow(g_fEpsilon, x)
# (src/libopenrave/libopenrave.h:140-142); the same values live in csrc/cc_epsilons.h for the device and the planner host
G_EPSILON = 1e-15
G_EPSILON_LINEAR = G_EPSILON ** 0.9              # g_fEpsilonLinear          3.16e-14
G_EPSILON_JOINT_LIMIT = G_EPSILON ** 0.8         # g_fEpsilonJointLimit      1e-12
G_EPSILON_EVAL_JOINT_LIMIT = G_EPSILON ** 0.65   # g_fEpsilonEvalJointLimit  1.78e-10

AO_Enabled = 1       # include/openrave/kinbody.h:3298
AO_ActiveDOFs = 2    # kinbody.h:3299


# ----------------------------------------------------------------------------- geometry
def _box_mesh(ex):
    """kinbodygeometry.cpp:538-569 — 8 corners, 12 triangles."""
    x, y, z = ex
    v = np.array([[x, y, z], [x, y, -z], [x, -y, z], [x, -y, -z],
                  [-x, y, z], [-x, y, -z], [-x, -y, z], [-x, -y, -z]], dtype=np.float64)
    idx = np.array([0, 2, 1, 1, 2, 3, 4, 5, 6, 5, 7, 6, 0, 1, 4, 1, 5, 4,
                    2, 6, 3, 3, 6, 7, 0, 4, 2, 2, 4, 6, 1, 3, 5, 3, 7, 5], dtype=np.int32).reshape(-1, 3)
    return v, idx


def _cylinder_mesh(radius, height, tess=1.0):
    """kinbodygeometry.cpp:213-271, 570-575 — z-axis cylinder, (int)(48*tess)+3 sides, 4 tris/side."""
    n = int(tess * 48.0) + 3
    hh = height * 0.5
    dtheta = 2 * rm.PI / float(n)
    verts = [(0.0, 0.0, 0.0 + hh), (0.0, 0.0, 0.0 - hh), (0.0 + radius, 0.0, 0.0 + hh), (0.0 + radius, 0.0, 0.0 - hh)]
    idx = []
    base = 0
    off = base + 4
    for i in range(1, n):
        ux = math.cos(dtheta * i)
        uy = math.sin(dtheta * i)
        verts.append((0.0 + ux * radius, 0.0 + uy * radius, 0.0 + hh))
        verts.append((0.0 + ux * radius, 0.0 + uy * radius, 0.0 - hh))
        idx += [base, off - 2, off, base + 1, off + 1, off - 1, off - 2, off - 1, off, off, off - 1, off + 1]
        off += 2
    idx += [base, off - 2, base + 2, base + 1, base + 3, off - 1, off - 2, off - 1, base + 2, base + 2, off - 1, base + 3]
    return np.array(verts, dtype=np.float64), np.array(idx, dtype=np.int32).reshape(-1, 3)


_ICO_X = 0.850650808352039932181540497063011072240401406   # sqrt(sqrt(5)+1)/sqrt(2*sqrt(5))
_ICO_Y = 0.525731112119133606025669084847876607285497935   # sqrt(2)/sqrt(5+sqrt(5))
_ICO_Z = 0.0


def _sphere_mesh(radius, levels=3):
    """kinbodygeometry.cpp:30-128 — unit icosahedron (GTS constants), each level splits every
    triangle in four with edge midpoints pushed back onto the sphere; then scaled by the radius."""
    X, Y, Z = _ICO_X, _ICO_Y, _ICO_Z
    verts = [(Z, X, -Y), (X, Y, Z), (Y, Z, -X), (Y, Z, X), (X, -Y, Z), (Z, X, Y),
             (-Y, Z, X), (Z, -X, -Y), (-X, Y, Z), (-Y, Z, -X), (-X, -Y, Z), (Z, -X, Y)]
    ind = [0, 1, 2, 1, 3, 4, 3, 5, 6, 2, 4, 7, 5, 6, 8, 2, 7, 9, 0, 5, 8, 7, 9, 10, 0, 1, 5, 7, 10, 11,
           1, 3, 5, 6, 10, 11, 3, 6, 11, 9, 10, 8, 3, 4, 11, 6, 8, 10, 4, 7, 11, 1, 2, 4, 0, 8, 9, 0, 2, 9]
    for i in range(0, 60, 3):
        v0, v1, v2 = verts[ind[i]], verts[ind[i + 1]], verts[ind[i + 2]]
        n = rm.cross((v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]), (v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]))
        if rm.dot3(v0, n) < 0:
            ind[i], ind[i + 1] = ind[i + 1], ind[i]
    while levels > 0:
        levels -= 1
        nverts = list(verts)
        nind = []
        edge = {}
        for i in range(0, len(ind), 3):
            tri = (ind[i], ind[i + 1], ind[i + 2])
            mids = []
            for j in range(3):
                a, b = tri[j], tri[(j + 1) % 3]
                if (a, b) not in edge:
                    edge[(a, b)] = edge[(b, a)] = len(nverts)
                    va, vb = verts[a], verts[b]
                    nverts.append(rm.normalize3((va[0] + vb[0], va[1] + vb[1], va[2] + vb[2])))
                mids.append(edge[(a, b)])
            nind += [tri[0], mids[0], mids[2], mids[0], tri[1], mids[1],
                     mids[2], mids[0], mids[1], mids[2], mids[1], tri[2]]
        verts, ind = nverts, nind
    v = np.array(verts, dtype=np.float64) * radius
    return v, np.array(ind, dtype=np.int32).reshape(-1, 3)


def transform_points(t, pts):
    """Transform * point for an [n,3] float64 array, written in the same association order as
    ravemath.quat_rotate so the vectorised result is bit-identical to the scalar one."""
    q = t[0]
    xx = 2 * q[1] * q[1]
    xy = 2 * q[1] * q[2]
    xz = 2 * q[1] * q[3]
    xw = 2 * q[1] * q[0]
    yy = 2 * q[2] * q[2]
    yz = 2 * q[2] * q[3]
    yw = 2 * q[2] * q[0]
    zz = 2 * q[3] * q[3]
    zw = 2 * q[3] * q[0]
    x, y, z = pts[:, 0], pts[:, 1], pts[:, 2]
    out = np.empty_like(pts)
    out[:, 0] = t[1][0] + ((1 - yy - zz) * x + (xy - zw) * y + (xz + yw) * z)
    out[:, 1] = t[1][1] + ((xy + zw) * x + (1 - xx - zz) * y + (yz - xw) * z)
    out[:, 2] = t[1][2] + ((xz - yw) * x + (yz + xw) * y + (1 - xx - yy) * z)
    return out


class Geometry(object):
    """KinBody::GeometryInfo subset: type, local transform, primitive data, collision mesh."""

    def __init__(self, gtype):
        self.type = gtype           # 'box' | 'cylinder' | 'sphere' | 'trimesh'
        self.t = rm.IDENT_T
        self.data = (0.0, 0.0, 0.0)  # half extents | (radius, height, _) | (radius,_,_)
        self.verts = np.zeros((0, 3), dtype=np.float64)
        self.idx = np.zeros((0, 3), dtype=np.int32)
        self.name = ""
        self.visible = True

    def init_collision_mesh(self, tess=1.0):
        if self.type == 'box':
            self.verts, self.idx = _box_mesh(self.data)
        elif self.type == 'cylinder':
            self.verts, self.idx = _cylinder_mesh(self.data[0], self.data[1], tess)
        elif self.type == 'sphere':
            self.verts, self.idx = _sphere_mesh(self.data[0], 3 + int(math.log(tess) / math.log(2.0)))


class Link(object):
    def __init__(self, body, name=""):
        self.body = body
        self.name = name
        self.index = -1
        self.static = False
        self.enabled = True
        self.t = rm.IDENT_T
        self.geometries = []
        self.verts = np.zeros((0, 3), dtype=np.float64)   # link-frame collision mesh (Link::_collision)
        self.idx = np.zeros((0, 3), dtype=np.int32)
        self.parent_links = []

    def update_collision(self):
        """Link::_Update, kinbodylink.cpp:1202-1228."""
        gs = self.geometries
        if len(gs) == 1 and gs[0].type == 'trimesh' and _transform_distance_fast(rm.IDENT_T, gs[0].t) <= G_EPSILON_LINEAR:
            self.verts, self.idx = gs[0].verts.copy(), gs[0].idx.copy()
            return
        vs, ids, off = [], [], 0
        for g in gs:
            vs.append(transform_points(g.t, g.verts))
            ids.append(g.idx + off)
            off += g.verts.shape[0]
        if vs:
            self.verts = np.concatenate(vs)
            self.idx = np.concatenate(ids).astype(np.int32)
        else:
            self.verts = np.zeros((0, 3), dtype=np.float64)
            self.idx = np.zeros((0, 3), dtype=np.int32)

    # openravepy-style accessors
    def GetName(self): return self.name
    def GetIndex(self): return self.index
    def GetParent(self): return self.body
    def IsEnabled(self): return self.enabled
    def IsStatic(self): return self.static
    def GetTransform(self): return self.t

    def Enable(self, enable):
        self.enabled = bool(enable)
        self.body._reset_collision_cache()


def _transform_distance_fast(t1, t2, frotweight=1.0, ftransweight=1.0):
    """geometry.h TransformDistanceFast: sqrt(|dtrans|^2) + min quat distance."""
    e1 = sum((a - b) * (a - b) for a, b in zip(t1[0], t2[0]))
    e2 = sum((a + b) * (a + b) for a, b in zip(t1[0], t2[0]))
    d = sum((a - b) * (a - b) for a, b in zip(t1[1], t2[1]))
    return math.sqrt(d) * ftransweight + frotweight * min(e1, e2)


# ----------------------------------------------------------------------------- mimic expressions
class MimicExpr(object):
    """A mimic_pos equation (an fparser expression over joint names in the reference,
    kinbodyjoint.cpp SetMimicEquations).  Plain arithmetic is parsed with Python's ast into a tiny
    evaluator; anything else of fparser's language is handed to the library's compiler for it;
    `affine()` extracts value = a*q[src] + b, the form the device chain program stores."""

    _ALLOWED = (ast.Expression, ast.BinOp, ast.UnaryOp, ast.Add, ast.Sub, ast.Mult, ast.Div, ast.Pow,
                ast.USub, ast.UAdd, ast.Name, ast.Load, ast.Constant)

    # names of fparser's functions (3rdparty/fparser-4.5.2/extrasrc/fptypes.hh:118-160) that the library's compiler knows
    _FUNCTIONS = frozenset("abs acos acosh asin asinh atan atan2 atanh cbrt ceil cos cosh cot csc exp exp2 floor hypot if int log "
                           "log10 log2 max min pow sec sin sinh sqrt tan tanh trunc SSSA SASA SASS".split())

    def __init__(self, text, names=None):
        """`names`: the joint names the equation may use.  Needed when a name is not an identifier ("dummy-1",
        "DarmSim/lowerTiltTopJoint" in COLLADA robots): names are then matched as the reference's substitution matches them,
        longest first (kinbodyjoint.cpp:2236-2244), before anything is parsed."""
        import re
        self.text = text.strip()
        self._code = None
        self._alias = None
        parse_text = self.text
        if names:
            pat = re.compile("|".join(re.escape(k) for k in sorted(set(names), key=len, reverse=True)))
            used = sorted(set(pat.findall(self.text)))
            self._alias = {k: "_v%d_" % i for i, k in enumerate(used)}
            parse_text = pat.sub(lambda m: self._alias[m.group(0)], self.text)
        try:
            self.tree = ast.parse(parse_text.replace("^", "**"), mode="eval")
            for node in ast.walk(self.tree):
                if not isinstance(node, self._ALLOWED):
                    raise ValueError("not plain arithmetic")
            found = sorted({n.id for n in ast.walk(self.tree) if isinstance(n, ast.Name)})
            if self._alias is not None:
                back = {v: k for k, v in self._alias.items()}
                if any(f not in back for f in found):
                    raise ValueError("unknown name")
                found = sorted(back[f] for f in found)
            self.varnames = found
            self._code = compile(self.tree, "<mimic>", "eval")
        except (ValueError, SyntaxError):
            # functions, comparisons, if(): the rest of fparser's language goes through the library's own compiler for it
            # (csrc/cc_mimic_expr.h, a host-side entry that needs no device)
            self.tree = None
            if self._alias is not None:
                self.varnames = sorted(self._alias)
            else:
                self.varnames = sorted(set(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", re.sub(r"\b0[xX][0-9a-fA-F]+|\b\d+\.?\d*([eE][-+]?\d+)?", " ", self.text)))
                                       - self._FUNCTIONS)
            try:
                self._library_eval([[0.0] * len(self.varnames)])
            except RuntimeError as e:
                raise ValueError("unsupported mimic expression %r: %s" % (text, e))

    def _library_eval(self, rows):
        import ctypes
        import numpy as np
        from .checker import load_library
        lib = load_library()
        names = (ctypes.c_char_p * max(len(self.varnames), 1))(*[v.encode() for v in self.varnames])
        v = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, max(len(self.varnames), 1))
        out = np.zeros(len(v))
        err = np.zeros(len(v), dtype=np.int32)
        if lib.cc_debug_mimic_eval(self.text.encode(), names, len(self.varnames), v.ctypes.data_as(ctypes.c_void_p), len(v),
                                   out.ctypes.data_as(ctypes.c_void_p), err.ctypes.data_as(ctypes.c_void_p)) != 0:
            raise RuntimeError(lib.cc_last_error().decode())
        return out                      # a row that failed to evaluate reads 0, as fparser's Eval returns

    def eval(self, values):
        if self._code is None:
            return float(self._library_eval([[float(values[v]) for v in self.varnames]])[0])
        return float(eval(self._code, {"__builtins__": {}}, self._env(values)))

    def _env(self, values):
        if self._alias is None:
            return dict(values)
        return {self._alias[k]: values[k] for k in self.varnames}

    def eval_many(self, values, n):
        """The same for n configurations at once: `values` maps joint names to arrays [n] (or scalars)."""
        import numpy as np
        if self._code is None:
            cols = [np.broadcast_to(np.asarray(values[v], dtype=np.float64), (n,)) for v in self.varnames]
            return self._library_eval(np.stack(cols, axis=1) if cols else np.zeros((n, 1)))
        return np.broadcast_to(np.asarray(eval(self._code, {"__builtins__": {}}, self._env(values)), dtype=np.float64), (n,)).copy()

    def affine(self):
        """Return (varname, a, b) if the expression is affine in exactly one joint, else None."""
        if len(self.varnames) != 1 or self._code is None:      # sampling proves nothing about min(), if(), floor() ...
            return None
        v = self.varnames[0]
        f0, f1 = self.eval({v: 0.0}), self.eval({v: 1.0})
        a, b = f1 - f0, f0
        for x in (-2.5, 0.37, 1.9, -0.71):
            fx = self.eval({v: x})
            if abs((a * x + b) - fx) > 1e-12 * max(1.0, abs(fx)):
                return None
        return v, a, b


# ----------------------------------------------------------------------------- joints
class Joint(object):
    def __init__(self, body, name="", jtype='hinge'):
        self.body = body
        self.name = name
        self.type = jtype              # 'hinge' (revolute) | 'slider' (prismatic)
        self.is_active = True          # XML enable= attribute; False => passive joint
        self.links = [None, None]      # _attachedbodies
        self.lower = -rm.PI if jtype == 'hinge' else -10.0
        self.upper = rm.PI if jtype == 'hinge' else 10.0
        self.offset = 0.0
        self.circular = False
        self.resolution = 0.02
        self.weight = 1.0
        self.mimic = None              # MimicExpr
        self.mimic_src = None          # resolved: list of (joint) the equation reads
        self.jointindex = -1
        self.dofindex = -1
        self.anchor = (0.0, 0.0, 0.0)
        self.axis = (0.0, 0.0, 1.0)    # _vaxes[0] (internal, after re-basing)
        self.t_left_nooffset = rm.IDENT_T
        self.t_right_nooffset = rm.IDENT_T
        self.t_left = rm.IDENT_T
        self.t_right = rm.IDENT_T
        self.static = None             # _nIsStatic

    def GetName(self): return self.name
    def GetDOFIndex(self): return self.dofindex
    def GetJointIndex(self): return self.jointindex
    def IsRevolute(self, iaxis=0): return self.type == 'hinge'
    def IsCircular(self, iaxis=0): return self.circular
    def IsMimic(self, iaxis=-1): return self.mimic is not None
    def GetLimits(self): return [self.lower], [self.upper]
    def GetHierarchyParentLink(self): return self.links[0]
    def GetHierarchyChildLink(self): return self.links[1]

    def compute_internal(self, link0, link1, t0, t1, anchor_raw, axis, currentvalue=None):
        """_ComputeJointInternalInformation for a 1-DOF hinge/slider (kinbodyjoint.cpp:1343-1493).
        t0/t1 are the link transforms to use (the reference reads them off the links)."""
        self.links = [link0, link1]
        vaxis = axis
        anchor = anchor_raw
        if self.links[1].static and self.links[0].index > 0 and not self.links[0].static:
            # keep the first body closer to the root (:1365-1374)
            tswap = rm.t_mul(rm.t_inv(t1), t0)
            r = rm.quat_rotate(tswap[0], vaxis)
            vaxis = (-r[0], -r[1], -r[2])
            anchor = rm.t_apply(tswap, anchor)
            self.links = [link1, link0]
            t0, t1 = t1, t0
        self.anchor = anchor
        trel = rm.t_mul(rm.t_inv(t0), t1)
        tl = (rm.IDENT_Q, anchor)
        tr = rm.t_mul((rm.IDENT_Q, (-anchor[0], -anchor[1], -anchor[2])), trel)
        if self.type == 'hinge':
            # re-base so the joint rotates about +z (:1425-1433)
            trot = (rm.quat_rotate_direction(vaxis, (0.0, 0.0, 1.0)), (0.0, 0.0, 0.0))
            tl = rm.t_mul(tl, rm.t_inv(trot))
            tr = rm.t_mul(trot, tr)
            vaxis = (0.0, 0.0, 1.0)
        self.axis = vaxis
        if self.type == 'hinge':
            toffset = (rm.quat_from_axis_angle(vaxis, self.offset), (0.0, 0.0, 0.0))
        else:
            toffset = (rm.IDENT_Q, (vaxis[0] * self.offset, vaxis[1] * self.offset, vaxis[2] * self.offset))
        self.t_left_nooffset = tl
        self.t_right_nooffset = tr
        self.t_left = rm.t_mul(tl, toffset)
        self.t_right = tr
        if currentvalue is not None:
            if self.type == 'hinge':
                toff = (rm.quat_from_axis_angle(vaxis, -currentvalue), (0.0, 0.0, 0.0))
            else:
                toff = (rm.IDENT_Q, (-vaxis[0] * currentvalue, -vaxis[1] * currentvalue, -vaxis[2] * currentvalue))
            self.t_left_nooffset = rm.t_mul(self.t_left_nooffset, toff)
            self.t_left = rm.t_mul(self.t_left, toff)
        self.static = None

    def is_static_now(self):
        """Joint::IsStatic before the cached flag exists (kinbodyjoint.cpp:810-848)."""
        if self.mimic is not None:
            srcs = self.mimic_src or []
            if srcs and all(j.is_static_now() for j in srcs):
                return True
            if srcs:
                return False
        if self.circular:
            return False
        return not (self.lower < self.upper)

    def compute_static(self):
        """_ComputeInternalStaticInformation (kinbodyjoint.cpp:1527-1551): a static joint folds its
        right transform into the left one."""
        if self.is_static_now():
            self.lower = 0.0
            self.upper = 0.0
            self.static = True
            self.t_left_nooffset = rm.t_mul(self.t_left_nooffset, self.t_right_nooffset)
            self.t_right_nooffset = rm.IDENT_T
            self.t_left = rm.t_mul(self.t_left, self.t_right)
            self.t_right = rm.IDENT_T
        else:
            self.static = False

    def IsStatic(self):
        return bool(self.static) if self.static is not None else self.is_static_now()


class Manipulator(object):
    def __init__(self, robot, name):
        self.robot = robot
        self.name = name
        self.base = None
        self.effector = None
        self.gripper_joint_names = []
        self.t_local = rm.IDENT_T

    def GetName(self): return self.name
    def GetEndEffector(self): return self.effector
    def GetBase(self): return self.base

    def GetArmIndices(self):
        """DOF indices of the active joints on the chain base -> effector (robotmanipulator.cpp)."""
        out = []
        for j in self.robot.get_chain_joints(self.base.index, self.effector.index):
            if j.dofindex >= 0 and not j.IsStatic() and j.mimic is None:
                out.append(j.dofindex)
        return out

    def GetTransform(self):
        return rm.t_mul(self.effector.t, self.t_local)

    def CheckEndEffectorCollision(self, report=None):
        """RobotBase::Manipulator::CheckEndEffectorCollision(report) (robotmanipulator.cpp:915-960): every link that all arm
        joints affect is checked against the environment; with CO_AllLinkCollisions the report aggregates over the links."""
        robot = self.robot
        env = robot.env
        cc = env.GetCollisionChecker()
        all_links = bool(cc.GetCollisionOptions() & 0x20)                    # CO_AllLinkCollisions
        arm = self.GetArmIndices()
        found, hit = [], False
        sub = type(report)() if report is not None else None
        for link in robot.links:
            if not all(robot.DoesAffect(robot.joints[d].jointindex, link.index) for d in arm):
                continue
            if env.CheckCollision(link, report=sub):
                hit = True
                if sub is not None:
                    found += list(sub.vLinkColliding)
                if not all_links:
                    break
        if report is not None:
            report.Reset(cc.GetCollisionOptions())
            for a, b in found:
                report.AddLinkCollision(a, b)
            if hit:
                report.minDistance = 0
        return hit


# ----------------------------------------------------------------------------- body
class KinBody(object):
    def __init__(self, env=None, name=""):
        self.env = env
        self.name = name
        self.links = []
        self.joints = []            # _vecjoints (active)
        self.passive_joints = []    # _vPassiveJoints
        self.forced_adjacent = set()
        self.make_joined_links_adjacent = True
        self.is_robot = False
        self.env_id = 0
        self._computed = False
        self.dof_values = []
        self.sorted_joints_all = []   # _vTopologicallySortedJointsAll
        self.adjacent = set()
        self.initial_link_transforms = []
        self._nonadjacent_cache = {}
        self.joints_affecting_links = None
        self.link_depths = []
        self.attached = []          # bodies rigidly attached through the environment (KinBody::GetAttached)
        self.grabbed = []           # KinBody::_vGrabbedBodies, in grab order
        self.uri = ""

    # --- openravepy-style accessors
    def GetName(self): return self.name
    def SetName(self, n): self.name = n
    def GetLinks(self): return self.links
    def GetJoints(self): return self.joints
    def GetPassiveJoints(self): return self.passive_joints
    def GetDOF(self): return len(self.joints)
    def IsRobot(self): return self.is_robot
    def GetEnv(self): return self.env
    def GetDOFValues(self): return list(self.dof_values)

    def GetLink(self, name):
        for l in self.links:
            if l.name == name:
                return l
        return None

    def GetJoint(self, name):
        for j in self.joints + self.passive_joints:
            if j.name == name:
                return j
        return None

    def GetDOFLimits(self):
        return [j.lower for j in self.joints], [j.upper for j in self.joints]

    def GetLinkTransformations(self):
        return [l.t for l in self.links]

    def GetTransform(self):
        return self.links[0].t

    def SetTransform(self, t):
        """KinBody::SetTransform: rigidly move every link so that link 0 lands on `t`."""
        tbaseinv = rm.t_inv(self.links[0].t)
        tapply = rm.t_mul(t, tbaseinv)
        for l in self.links:
            l.t = rm.t_mul(tapply, l.t)
        self._notify_transform_changed()

    def SetLinkTransformations(self, ts):
        for l, t in zip(self.links, ts):
            l.t = t
        self._notify_transform_changed()

    def Enable(self, enable):
        for l in self.links:
            l.enabled = bool(enable)
        self._reset_collision_cache()

    def IsEnabled(self):
        return any(l.enabled for l in self.links)

    def _notify_transform_changed(self):
        self._update_grabbed_bodies()
        if self.env is not None:
            self.env._body_changed(self, "pose")

    # --- grabbing (src/libopenrave/kinbodygrab.cpp) -------------------------------------------------
    def rigidly_attached_links(self, link):
        """Link::GetRigidlyAttachedLinks: the links joined to `link` through static joints (itself included)."""
        group = {link.index}
        changed = True
        while changed:
            changed = False
            for j in self.all_joints():
                if j.IsStatic():
                    a, b = j.links[0].index, j.links[1].index
                    if (a in group) != (b in group):
                        group.update((a, b))
                        changed = True
        return [self.links[i] for i in sorted(group)]

    def Grab(self, body, link, links_to_ignore=()):
        """KinBody::Grab(pGrabbedBody, pGrabbingLink, setGrabberLinksToIgnore) (kinbodygrab.cpp:204-330): the body follows `link`
        from now on; Grabbed::ComputeListNonCollidingLinks (:63-200) is evaluated here, at the grab-time state, with every
        link of both bodies enabled and the body not yet attached."""
        if body is self or link.body is not self:
            raise ValueError("Grab: bad body / link")
        self.Release(body)
        cc = self.env.GetCollisionChecker() if self.env is not None else None
        if cc is None:
            raise RuntimeError("Grab needs a collision checker (the non-colliding link list is a collision query)")
        g = Grabbed(body, link, set(int(i) for i in links_to_ignore))
        saved = [(l, l.enabled) for l in self.links + body.links] + [(l, l.enabled) for o in self.grabbed for l in o.body.links]
        for l, _ in saved:
            l.enabled = True
        for b in [self, body] + [o.body for o in self.grabbed]:
            b._reset_collision_cache()
        try:
            rigid = set(l.index for l in self.rigidly_attached_links(link))
            for l in self.links:
                if l.index in rigid or l.index in g.ignore:
                    continue
                if not cc.CheckCollision(l, body):
                    g.noncolliding.append(l)
            for o in self.grabbed:
                if o.link.index in rigid:       # constant relative pose: never checked against each other
                    continue
                for l in o.body.links:
                    if not cc.CheckCollision(l, body):
                        g.noncolliding.append(l)
        finally:
            for l, e in saved:
                l.enabled = e
            for b in [self, body] + [o.body for o in self.grabbed]:
                b._reset_collision_cache()
        self.grabbed.append(g)
        for a, b in ((self, body), (body, self)):
            if b not in a.attached:
                a.attached.append(b)
        self._grab_changed()
        return True

    def Release(self, body):
        for g in list(self.grabbed):
            if g.body is body:
                self.grabbed.remove(g)
                if body in self.attached:
                    self.attached.remove(body)
                if self in body.attached:
                    body.attached.remove(self)
                for o in self.grabbed:            # its links leave the other grabbed bodies' lists
                    o.noncolliding = [l for l in o.noncolliding if l.body is not body]
                self._grab_changed()

    def ReleaseAllGrabbed(self):
        for g in list(self.grabbed):
            self.Release(g.body)

    def GetGrabbed(self):
        return [g.body for g in self.grabbed]

    def IsGrabbing(self, body):
        return any(g.body is body for g in self.grabbed)

    def IsAttached(self, body):
        """KinBody::IsAttached: transitive over attachments."""
        seen, todo = {id(self)}, [self]
        while todo:
            for o in todo.pop().attached:
                if o is body:
                    return True
                if id(o) not in seen:
                    seen.add(id(o))
                    todo.append(o)
        return body is self

    def _grab_changed(self):
        self.grab_epoch = getattr(self, "grab_epoch", 0) + 1
        if self.env is not None:
            self.env._body_changed(self)

    def _update_grabbed_bodies(self):
        """KinBody::_UpdateGrabbedBodies: every grabbed link = grabbing link * relative pose."""
        for g in self.grabbed:
            for l, rel in zip(g.body.links, g.t_rel):
                l.t = rm.t_mul(g.link.t, rel)
            g.body._update_grabbed_bodies()
            if g.body.env is not None:
                g.body.env._body_changed(g.body, "pose")

    def GetQueryLinkTransformations(self):
        """The body's link transforms followed by those of its grabbed bodies in grab order: the row layout of the
        checkers' pose-mode entries (cc_check_poses) when bodies are grabbed."""
        ts = [l.t for l in self.links]
        for g in self.grabbed:
            ts += [l.t for l in g.body.links]
        return ts

    def _reset_collision_cache(self):
        self._nonadjacent_cache = {k: v for k, v in self._nonadjacent_cache.items() if k[1] == 0}
        if self.env is not None:
            self.env._body_changed(self)

    def all_joints(self):
        return self.joints + self.passive_joints

    # --- hierarchy -----------------------------------------------------------------------------
    def compute_internal_information(self):
        """KinBody::_ComputeInternalInformation for tree-structured bodies (kinbody.cpp:4140-5050)."""
        nl = len(self.links)
        for i, l in enumerate(self.links):
            l.index = i
            if not hasattr(l, "raw_t"):
                l.raw_t = l.t            # pose as parsed (serialize.py stores the pre-FK state)
        for i, j in enumerate(self.joints):
            j.jointindex = i
            j.dofindex = i
        for j in self.joints + self.passive_joints:
            if getattr(j, "raw", None) is None:
                j.raw = dict(l0=j.links[0].index, l1=j.links[1].index, lower=j.lower, upper=j.upper, axis=j.axis,
                             anchor=j.anchor, tl=j.t_left_nooffset, tr=j.t_right_nooffset, tlo=j.t_left, tro=j.t_right)
        for j in self.passive_joints:
            j.jointindex = -1
            j.dofindex = -1
        alljoints = self.all_joints()
        byname = {j.name: j for j in alljoints}
        for j in alljoints:
            if j.mimic is not None:
                j.mimic_src = []
                for vn in j.mimic.varnames:
                    if vn not in byname:
                        raise ValueError("mimic joint %s references unknown joint %s" % (j.name, vn))
                    j.mimic_src.append(byname[vn])
        # undirected link graph, neighbours sorted (kinbody.cpp:4447-4460)
        nbr = [[] for _ in range(nl)]
        edge_joint = {}
        for gi, j in enumerate(alljoints):
            a, b = j.links[0].index, j.links[1].index
            nbr[a].append(b)
            nbr[b].append(a)
            if (a, b) in edge_joint or (b, a) in edge_joint:
                raise NotImplementedError("%s: parallel joints / closed loops are not supported" % self.name)
            edge_joint[(a, b)] = gi
            edge_joint[(b, a)] = gi
        if len(alljoints) > nl - 1 and nl > 0:
            raise NotImplementedError("%s: closed kinematic loops are not supported" % self.name)
        for a in nbr:
            a.sort()
        # depths = minimum path length to the root / a static link (:4525-4552)
        depth = [-1] * nl
        parent = [-1] * nl
        roots = [0] + [i for i, l in enumerate(self.links) if l.static and i != 0]
        for r in roots:
            depth[r] = 0
        frontier = [0]
        seen = {0}
        order_links = [0]
        while frontier:
            nxt = []
            for u in frontier:
                for v in nbr[u]:
                    if v not in seen:
                        seen.add(v)
                        parent[v] = u
                        if depth[v] < 0:
                            depth[v] = depth[u] + 1
                        nxt.append(v)
                        order_links.append(v)
            frontier = nxt
        self.link_depths = depth
        self._parent_link = parent
        self._edge_joint = edge_joint
        self._alljoints = alljoints
        for i, l in enumerate(self.links):
            l.parent_links = [parent[i]] if parent[i] >= 0 else []
        # order: parents before children; a mimic joint after the passive joints it reads
        sorted_gi = []
        for v in order_links[1:]:
            sorted_gi.append(edge_joint[(parent[v], v)])
        # joints whose links are not connected to link 0 keep declaration order at the end
        for gi in range(len(alljoints)):
            if gi not in sorted_gi:
                sorted_gi.append(gi)
        changed = True
        while changed:
            changed = False
            for pos, gi in enumerate(sorted_gi):
                j = alljoints[gi]
                if j.mimic is not None:
                    for src in j.mimic_src:
                        if src.dofindex < 0:
                            sp = sorted_gi.index(alljoints.index(src))
                            if sp > pos:
                                sorted_gi.insert(pos, sorted_gi.pop(sp))
                                changed = True
        # parent link of each joint must be its first attached link; swap otherwise (:4713-4784)
        for gi in sorted_gi:
            j = alljoints[gi]
            a, b = j.links[0].index, j.links[1].index
            want_parent = a
            if (j.links[0].static or j.links[1].static) and (a != 0 and b != 0):
                raise NotImplementedError("%s: joints on static non-root links are not supported" % self.name)
            if parent[a] == b:   # declared child-first: the tree parent must be the first attached link
                want_parent = b
            if want_parent != a:
                tswap = rm.t_inv(j.t_right_nooffset)
                r = rm.quat_rotate(tswap[0], j.axis)
                vaxis = (-r[0], -r[1], -r[2])
                t_first = rm.IDENT_T
                t_second = rm.t_mul(j.t_left_nooffset, j.t_right_nooffset)
                j.compute_internal(j.links[1], j.links[0], t_second, t_first, tswap[1], vaxis, None)
            j.compute_static()
        self.sorted_joints_all = [alljoints[gi] for gi in sorted_gi]
        # joints affecting links (:4787-4818): every active joint on the path root -> link, plus the
        # DOFs a mimic joint on that path reads
        nj = len(self.joints)
        aff = [[0] * nl for _ in range(nj)]
        for i in range(nl):
            v = i
            while parent[v] >= 0:
                gi = edge_joint[(parent[v], v)]
                j = alljoints[gi]
                if gi < nj:
                    aff[gi][i] = 1
                if j.mimic is not None:
                    for src in self._mimic_dof_sources(j):
                        aff[src.jointindex][i] = 1
                v = parent[v]
        self.joints_affecting_links = aff
        # adjacency (:4940-5005)
        adj = set(self.forced_adjacent)
        for l0 in self.links:
            if len(l0.geometries) == 0:
                for l1 in self.links:
                    if l0.index != l1.index:
                        adj.add(_pair(l0.index, l1.index))
        if self.make_joined_links_adjacent:
            for j in alljoints:
                adj.add(_pair(j.links[0].index, j.links[1].index))
            for i0 in range(nl - 1):
                for i1 in range(i0 + 1, nl):
                    chain = self.get_chain_joints(i0, i1)
                    numstatic = sum(1 for j in chain if j.IsStatic())
                    if numstatic + 1 >= len(chain):
                        adj.add((i0, i1))
        self.adjacent = adj
        self._nonadjacent_cache = {}
        self._computed = True
        # SetDOFValues once so mimic joints take effect; this defines the initial pose (:5026-5048)
        if not self.dof_values or len(self.dof_values) != nj:
            self.dof_values = [0.0] * nj
        self.SetDOFValues(self.dof_values)
        self.initial_link_transforms = [l.t for l in self.links]

    def _mimic_dof_sources(self, j, _seen=None):
        out = []
        for src in j.mimic_src or []:
            if src.dofindex >= 0:
                out.append(src)
            elif src.mimic is not None:
                out += self._mimic_dof_sources(src)
        return out

    def get_chain_joints(self, i0, i1):
        """Joints on the tree path between two links (KinBody::GetChain, kinbody.cpp:2559-2578)."""
        parent = self._parent_link
        anc0 = []
        v = i0
        while v >= 0:
            anc0.append(v)
            v = parent[v]
        pos = {v: k for k, v in enumerate(anc0)}
        path1 = []
        v = i1
        while v >= 0 and v not in pos:
            path1.append(v)
            v = parent[v]
        if v < 0:
            return []   # disconnected
        joints = []
        for k in range(pos[v]):
            joints.append(self._alljoints[self._edge_joint[(parent[anc0[k]], anc0[k])]])
        for u in reversed(path1):
            joints.append(self._alljoints[self._edge_joint[(parent[u], u)]])
        return joints

    def DoesAffect(self, jointindex, linkindex):
        return self.joints_affecting_links[jointindex][linkindex]

    def IsDOFInChain(self, l0, l1, dofindex):
        return (self.DoesAffect(dofindex, l0) == 0) != (self.DoesAffect(dofindex, l1) == 0)

    def AreAdjacentLinks(self, i0, i1):
        return _pair(i0, i1) in self.adjacent

    # --- forward kinematics -------------------------------------------------------------------
    def SetDOFValues(self, values, dofindices=None, checklimits=True):
        """KinBody::SetDOFValues (kinbody.cpp:2121-2485), 1-DOF hinge/slider joints."""
        vals = list(self.dof_values) if self.dof_values else [0.0] * len(self.joints)
        if dofindices is not None:
            for v, i in zip(values, dofindices):
                if not math.isnan(v):
                    vals[i] = float(v)
        else:
            for i in range(len(vals)):
                if not math.isnan(values[i]):
                    vals[i] = float(values[i])
        if checklimits:
            for j in self.joints:
                if j.circular:
                    continue
                p = vals[j.dofindex]
                if p < j.lower:
                    vals[j.dofindex] = j.lower
                elif p > j.upper:
                    vals[j.dofindex] = j.upper
        self.dof_values = vals
        ts = fk_transforms(self, vals, self.links[0].t, checklimits)
        for l, t in zip(self.links, ts):
            l.t = t
        self._notify_transform_changed()

    # --- non-adjacent pairs ---------------------------------------------------------------------
    def GetNonAdjacentLinks(self, adjacentoptions=0, checker=None):
        """kinbody.cpp:5532-5600: pairs (i<j) that are not adjacent and do not collide at the initial
        pose, sorted by CompareNonAdjacentFarthest; packed i | j<<16."""
        cc = checker if checker is not None else (self.env.GetCollisionChecker() if self.env else None)
        ck = id(cc)
        if (ck, 0) not in self._nonadjacent_cache:
            if cc is None:
                raise RuntimeError("GetNonAdjacentLinks needs a collision checker (it is itself a collision query)")
            cands = []
            nl = len(self.links)
            for i0 in range(nl):
                for i1 in range(i0 + 1, nl):
                    if not self.AreAdjacentLinks(i0, i1):
                        cands.append((i0, i1))
            hits = cc._check_link_pairs_at(self, self.initial_link_transforms, cands)
            base = [i0 | (i1 << 16) for (i0, i1), h in zip(cands, hits) if not h]
            self._nonadjacent_cache[(ck, 0)] = sort_nonadjacent(base)
        if (ck, adjacentoptions) not in self._nonadjacent_cache:
            cur = self._nonadjacent_cache[(ck, 0)]
            if adjacentoptions & AO_ActiveDOFs:
                active = self.GetActiveDOFIndices() if self.is_robot else list(range(self.GetDOF()))
                cur = [p for p in cur if any(self.IsDOFInChain(p & 0xffff, p >> 16, d) for d in active)]
            if adjacentoptions & AO_Enabled:
                cur = [p for p in cur if self.links[p & 0xffff].enabled and self.links[p >> 16].enabled]
            self._nonadjacent_cache[(ck, adjacentoptions)] = sort_nonadjacent(cur)
        return self._nonadjacent_cache[(ck, adjacentoptions)]

    def CheckSelfCollision(self, report=None, collisionchecker=None):
        """KinBody::CheckSelfCollision (kinbodycollision.cpp:21-223): the standalone check, then per grabbed body its
        non-colliding links x its links, its own self collision, and the later grabbed bodies link by link.  This is the
        literal per-call loop; the batched entry evaluates the same pair set in one pass (cc_query_create)."""
        cc = collisionchecker if collisionchecker is not None else self.env.GetCollisionChecker()
        if cc.CheckStandaloneSelfCollision(self, report):
            return True
        for i1, g1 in enumerate(self.grabbed):
            if not g1.body.IsEnabled():
                continue
            for rl in g1.noncolliding:
                for gl in g1.body.links:
                    if cc.CheckCollision(rl, gl, report=report):
                        return True
            if g1.body.CheckSelfCollision(report, cc):
                return True
            for g2 in self.grabbed[i1 + 1:]:
                if not g2.body.IsEnabled():
                    continue
                for l2 in g2.body.links:
                    if not l2.enabled or any(l2 is x for x in g1.noncolliding):
                        continue
                    for l1 in g1.body.links:
                        if l1.enabled and any(l1 is x for x in g2.noncolliding):
                            if cc.CheckCollision(l1, l2, report=report):
                                return True
        return False


class Grabbed(object):
    """One entry of KinBody::_vGrabbedBodies (kinbodygrab.h)."""

    def __init__(self, body, link, ignore):
        self.body = body
        self.link = link
        self.ignore = ignore
        tinv = rm.t_inv(link.t)
        self.t_rel = [rm.t_mul(tinv, l.t) for l in body.links]
        self.noncolliding = []          # _listNonCollidingLinksWhenGrabbed: links of the grabber and of its other grabbed bodies


def _pair(a, b):
    return (a, b) if a < b else (b, a)


def sort_nonadjacent(pairs):
    """CompareNonAdjacentFarthest (kinbody.cpp:5512-5530): larger j-i first, then larger j, then larger i."""
    def key(p):
        i, j = p & 0xffff, (p >> 16) & 0xffff
        return (-(j - i), -j, -i)
    return sorted(pairs, key=key)


def fk_transforms(body, vals, tbase, checklimits=True):
    """The joint loop of SetDOFValues (kinbody.cpp:2243-2481) as a pure function:
    DOF values (already clamped) -> list of link transforms.  Link 0 keeps `tbase`."""
    nl = len(body.links)
    ts = [l.t for l in body.links]
    ts[0] = tbase
    computed = [False] * nl
    computed[0] = True
    passive_vals = {}
    for j in body.sorted_joints_all:
        parent, child = j.links[0], j.links[1]
        tparent = ts[parent.index]
        if j.IsStatic():
            ts[child.index] = rm.t_mul(tparent, j.t_left_nooffset)
            computed[child.index] = True
            continue
        if j.mimic is not None:
            env = {}
            for src in j.mimic_src:
                env[src.name] = vals[src.dofindex] if src.dofindex >= 0 else passive_vals.get(src.name, 0.0)
            ev = j.mimic.eval(env)
            v = ev
            if not j.circular:
                # kinbody.cpp:2326-2386: snap within g_fEpsilonJointLimit, otherwise clamp when checking limits
                if ev < j.lower:
                    if ev >= j.lower - G_EPSILON_JOINT_LIMIT:
                        v = j.lower
                    elif checklimits and ev < j.lower - G_EPSILON_EVAL_JOINT_LIMIT:
                        v = j.lower
                elif ev > j.upper:
                    if ev <= j.upper + G_EPSILON_JOINT_LIMIT:
                        v = j.upper
                    elif checklimits and ev > j.upper + G_EPSILON_EVAL_JOINT_LIMIT:
                        v = j.upper
            value = v
            if j.dofindex < 0:
                passive_vals[j.name] = value
        elif j.dofindex >= 0:
            value = vals[j.dofindex]
        else:
            value = passive_vals.get(j.name, 0.0)
            if not j.circular:
                value = min(max(value, j.lower), j.upper)
        if computed[child.index]:
            continue
        if j.type == 'hinge':
            tj = (rm.quat_from_axis_angle(j.axis, value), (0.0, 0.0, 0.0))
        else:
            tj = (rm.IDENT_Q, (j.axis[0] * value, j.axis[1] * value, j.axis[2] * value))
        tlocal = rm.t_mul(rm.t_mul(j.t_left_nooffset, tj), j.t_right_nooffset)
        ts[child.index] = rm.t_mul(tparent, tlocal)
        computed[child.index] = True
    return ts


def external_mimic_values(body, vals, externals, checklimits=True):
    """What a caller of the batched entry supplies for mimic joints whose equation the device cannot evaluate (CC_MIMIC_EXTERNAL):
    their RAW equation values for n full DOF vectors `vals` [n, ndof] (already clamped like SetDOFValues clamps them,
    kinbody.cpp:2149-2224), evaluated the way SetDOFValues' joint loop does (kinbody.cpp:2308-2386): a source that is itself a
    passive mimic joint contributes its clamped value.  Returns [n, len(externals)]."""
    import numpy as np
    n = vals.shape[0]
    out = np.zeros((n, len(externals)), dtype=np.float64)
    passive_vals = {}
    for j in body.sorted_joints_all:
        if j.IsStatic() or j.mimic is None:
            continue
        env = {}
        for src in j.mimic_src:
            env[src.name] = vals[:, src.dofindex] if src.dofindex >= 0 else passive_vals.get(src.name, np.zeros(n))
        ev = j.mimic.eval_many(env, n)
        for k, e in enumerate(externals):
            if e is j:
                out[:, k] = ev
        v = ev.copy()
        if not j.circular:
            lo = (ev < j.lower) & ((ev >= j.lower - G_EPSILON_JOINT_LIMIT) | (checklimits & (ev < j.lower - G_EPSILON_EVAL_JOINT_LIMIT)))
            hi = (ev > j.upper) & ((ev <= j.upper + G_EPSILON_JOINT_LIMIT) | (checklimits & (ev > j.upper + G_EPSILON_EVAL_JOINT_LIMIT)))
            v[lo] = j.lower
            v[hi] = j.upper
        if j.dofindex < 0:
            passive_vals[j.name] = v
    return out


class Robot(KinBody):
    def __init__(self, env=None, name=""):
        KinBody.__init__(self, env, name)
        self.is_robot = True
        self.manipulators = []
        self.active_dofs = None
        self.active_manip = 0

    def GetManipulators(self): return self.manipulators

    def GetActiveManipulator(self):
        return self.manipulators[self.active_manip] if self.manipulators else None

    def SetActiveDOFs(self, indices, affine=0, rotationaxis=(0.0, 0.0, 1.0)):
        """RobotBase::SetActiveDOFs(dofindices, affine, rotationaxis) (robot.cpp:1130-1190): joint DOFs plus OpenRAVE::DOFAffine
        bits; an active-DOF vector is the joint values followed by X | Y | Z and one rotation parameterisation."""
        self.active_dofs = [int(i) for i in indices]
        self.affine_dofs = int(affine)
        self.affine_axis = tuple(float(x) for x in rotationaxis)
        self.affine_quat_start = (1.0, 0.0, 0.0, 0.0)                      # _vRotationQuatLimitStart (robot.cpp:826)
        self._nonadjacent_cache = {k: v for k, v in self._nonadjacent_cache.items() if not (k[1] & AO_ActiveDOFs)}
        if self.env is not None:
            self.env._body_changed(self)

    def GetActiveDOFIndices(self):
        return list(range(self.GetDOF())) if self.active_dofs is None else list(self.active_dofs)

    def GetAffineDOF(self):
        return int(getattr(self, "affine_dofs", 0))

    def GetActiveDOF(self):
        from .flatten import affine_count
        return len(self.GetActiveDOFIndices()) + affine_count(self.GetAffineDOF())

    def SetActiveDOFValues(self, values):
        """RobotBase::SetActiveDOFValues (robot.cpp:1193-1225): the affine values overwrite components of the base pose
        (RaveGetTransformFromAffineDOFValues, libopenrave.cpp:1868-1924), then SetDOFValues(joints, transform)."""
        idx = self.GetActiveDOFIndices()
        m = self.GetAffineDOF()
        if not m:
            self.SetDOFValues(values, idx)
            return
        a = [float(v) for v in values[len(idx):]]
        q, t = self.links[0].t
        q, t = list(q), list(t)
        k = 0
        for bit, comp in ((1, 0), (2, 1), (4, 2)):
            if m & bit:
                t[comp] = a[k]
                k += 1
        if m & 8:
            s = math.sin(a[k] * 0.5)
            q = [math.cos(a[k] * 0.5), self.affine_axis[0] * s, self.affine_axis[1] * s, self.affine_axis[2] * s]
        elif m & 16:
            x, y, z = a[k:k + 3]
            ang = math.sqrt(x * x + y * y + z * z)
            if ang > 0:
                f = math.sin(0.5 * ang) / ang
                q = [math.cos(0.5 * ang), f * x, f * y, f * z]
            else:
                q = [1.0, 0.0, 0.0, 0.0]
        elif m & 32:
            r = rm.normalize4(tuple(a[k:k + 4]))
            q = list(rm.quat_mul_raw(self.affine_quat_start, r))
        self.SetTransform((tuple(q), tuple(t)))
        if idx:
            self.SetDOFValues(list(values[:len(idx)]), idx)

    def GetActiveDOFLimits(self):
        lo, hi = self.GetDOFLimits()
        idx = self.GetActiveDOFIndices()
        return [lo[i] for i in idx], [hi[i] for i in idx]
