"""Flatten a computed KinBody into the plain arrays the C ABI takes (include/cudacollision.h).

This is what the OpenRAVE-side plugin does with a live KinBody: link meshes
from Link::GetCollisionData(), poses from GetLinkTransformations(), the FK chain from
GetDependencyOrderedJointsAll() + Joint::GetInternalHierarchy{Left,Right}Transform/Axis, the
DoesAffect table and GetNonAdjacentLinks(0).
"""
import ctypes

import numpy as np

from . import ravemath as rm


class cc_joint(ctypes.Structure):
    """Mirror of `struct cc_joint` in include/cudacollision.h."""
    _fields_ = [("type", ctypes.c_int32), ("parent_link", ctypes.c_int32), ("child_link", ctypes.c_int32),
                ("dof_index", ctypes.c_int32), ("mimic_dof", ctypes.c_int32), ("is_static", ctypes.c_int32),
                ("is_circular", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("mimic_a", ctypes.c_double), ("mimic_b", ctypes.c_double),
                ("lower", ctypes.c_double), ("upper", ctypes.c_double),
                ("left", ctypes.c_double * 8), ("right", ctypes.c_double * 8), ("axis", ctypes.c_double * 4)]


CC_JOINT_REVOLUTE, CC_JOINT_PRISMATIC = 1, 2


def transforms8(ts):
    """list of (quat, trans) -> float64 [n, 8]."""
    out = np.zeros((len(ts), 8), dtype=np.float64)
    for i, t in enumerate(ts):
        out[i, 0:4] = t[0]
        out[i, 4:7] = t[1]
    return out


def transforms_from8(arr):
    return [((float(r[0]), float(r[1]), float(r[2]), float(r[3])), (float(r[4]), float(r[5]), float(r[6]))) for r in arr]


def link_mesh_arrays(link):
    v = np.ascontiguousarray(link.verts, dtype=np.float64)
    i = np.ascontiguousarray(link.idx, dtype=np.int32)
    return v, i


CC_MIMIC_EXTERNAL = -2


def joint_table(body, externals=None):
    """The FK chain program: one cc_joint per entry of the dependency-ordered joint list.  A mimic equation that is not
    a * q[dof] + b can only go to a caller that evaluates it per configuration (CC_MIMIC_EXTERNAL): pass a list as `externals`
    and it receives those joints in column order; without it such a body raises NotImplementedError."""
    js = body.sorted_joints_all
    arr = (cc_joint * max(1, len(js)))()
    for k, j in enumerate(js):
        c = arr[k]
        c.type = CC_JOINT_REVOLUTE if j.type == 'hinge' else CC_JOINT_PRISMATIC
        c.parent_link = j.links[0].index
        c.child_link = j.links[1].index
        c.dof_index = j.dofindex
        c.mimic_dof = -1
        c.mimic_a, c.mimic_b = 0.0, 0.0
        if j.mimic is not None:
            aff = j.mimic.affine()
            src = body.GetJoint(aff[0]) if aff is not None else None
            if aff is None or src is None or src.dofindex < 0:
                if externals is None or j.IsStatic():
                    if aff is None:
                        raise NotImplementedError("mimic joint %s: only affine single-DOF equations run on the device (%r)"
                                                  % (j.name, j.mimic.text))
                    raise NotImplementedError("mimic joint %s must read an active DOF" % j.name)
                c.mimic_dof = CC_MIMIC_EXTERNAL
                c.reserved = len(externals)
                externals.append(j)
            else:
                c.mimic_dof = src.dofindex
                c.mimic_a, c.mimic_b = aff[1], aff[2]
            c.dof_index = -1
        c.is_static = 1 if j.IsStatic() else 0
        c.is_circular = 1 if j.circular else 0
        c.lower, c.upper = j.lower, j.upper
        for i, v in enumerate(rm.t_to_array8(j.t_left_nooffset)):
            c.left[i] = v
        for i, v in enumerate(rm.t_to_array8(j.t_right_nooffset)):
            c.right[i] = v
        c.axis[0], c.axis[1], c.axis[2], c.axis[3] = j.axis[0], j.axis[1], j.axis[2], 0.0
    return arr, len(js)


def affects_table(body):
    nd, nl = body.GetDOF(), len(body.links)
    out = np.zeros((max(nd, 1), nl), dtype=np.uint8)
    for d in range(nd):
        for l in range(nl):
            out[d, l] = 1 if body.DoesAffect(d, l) else 0
    return out


def grab_tables(body, body_id):
    """KinBody::_vGrabbedBodies -> the arguments of cc_body_set_grabbed, one tuple per grabbed body in grab order:
    (grabbed body id, grabbing link index, rel [nlinks, 8] float64, noncolliding [(body id, link index)] int32 [m, 2]).
    `body_id(kinbody)` maps a KinBody to the checker's body id."""
    out = []
    for g in getattr(body, "grabbed", []):
        rel = transforms8(g.t_rel)
        nc = np.array([(body_id(l.body), l.index) for l in g.noncolliding], dtype=np.int32).reshape(-1, 2)
        out.append((body_id(g.body), g.link.index, rel, nc))
    return out


DOF_X, DOF_Y, DOF_Z, DOF_RotationAxis, DOF_Rotation3D, DOF_RotationQuat = 1, 2, 4, 8, 16, 32      # OpenRAVE::DOFAffine (openrave.h:2540-2551)


def affine_count(mask):
    """RaveGetAffineDOF: how many values the affine DOFs of `mask` add to an active-DOF vector."""
    return sum(1 for b in (DOF_X, DOF_Y, DOF_Z, DOF_RotationAxis) if mask & b) + (3 if mask & DOF_Rotation3D else 0) + \
        (4 if mask & DOF_RotationQuat else 0)
