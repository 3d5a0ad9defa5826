"""Scalar SE(3) arithmetic in IEEE double, following OpenRAVE's geometry.h.

Host-side helper for the scene loader and the Python mirror of KinBody.  Pure Python
floats are IEEE doubles, so every function here reproduces the reference's `dReal`
arithmetic operation by operation (same association order, same normalisation rule).

Conventions (reference: include/openrave/geometry.h)
  * quaternion = (w, x, y, z) with the scalar part FIRST (geometry.h:524, identity :398-400)
  * transform  = (quat, trans) where trans = (x, y, z)
"""
import math

EPS = 2.220446049250313e-16  # std::numeric_limits<double>::epsilon()
PI = 3.14159265358979323846

IDENT_Q = (1.0, 0.0, 0.0, 0.0)
IDENT_T = (IDENT_Q, (0.0, 0.0, 0.0))


def normalize4(q):
    """geometry.h:255-264 — divide only when |q|^2 is off by more than machine eps."""
    f = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]
    if f < 1.0 - EPS or f > 1.0 + EPS:
        f = math.sqrt(f)
        return (q[0] / f, q[1] / f, q[2] / f, q[3] / f)
    return q


def normalize3(v):
    """geometry.h:265-273."""
    f = v[0] * v[0] + v[1] * v[1] + v[2] * v[2]
    if f < 1.0 - EPS or f > 1.0 + EPS:
        f = math.sqrt(f)
        return (v[0] / f, v[1] / f, v[2] / f)
    return v


def quat_from_axis_angle(axis, angle):
    """geometry.h:673-682 — divides by |axis|; a zero axis gives the identity."""
    axislen = math.sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2])
    if axislen == 0:
        return IDENT_Q
    angle *= 0.5
    sang = math.sin(angle) / axislen
    return (math.cos(angle), axis[0] * sang, axis[1] * sang, axis[2] * sang)


def quat_rotate(q, r):
    """RaveTransform::rotate(vector), geometry.h:430-446."""
    xx = 2 * q[1] * q[1]
    xy = 2 * q[1] * q[2]
    xz = 2 * q[1] * q[3]
    xw = 2 * q[1] * q[0]
    yy = 2 * q[2] * q[2]
    yz = 2 * q[2] * q[3]
    yw = 2 * q[2] * q[0]
    zz = 2 * q[3] * q[3]
    zw = 2 * q[3] * q[0]
    return ((1 - yy - zz) * r[0] + (xy - zw) * r[1] + (xz + yw) * r[2],
            (xy + zw) * r[0] + (1 - xx - zz) * r[1] + (yz - xw) * r[2],
            (xz - yw) * r[0] + (yz + xw) * r[1] + (1 - xx - yy) * r[2])


def quat_mul_raw(a, b):
    """Hamilton product without normalisation (geometry.h:1096-1105)."""
    return (a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3],
            a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1])


def t_apply(t, v):
    """Transform * vector: trans + rotate(v) (geometry.h:425-427)."""
    r = quat_rotate(t[0], v)
    return (t[1][0] + r[0], t[1][1] + r[1], t[1][2] + r[2])


def t_mul(a, b):
    """Transform * Transform with the per-compose normalize4 (geometry.h:466-480)."""
    return (normalize4(quat_mul_raw(a[0], b[0])), t_apply(a, b[1]))


def t_inv(t):
    """geometry.h:500-509."""
    qi = (t[0][0], -t[0][1], -t[0][2], -t[0][3])
    r = quat_rotate(qi, t[1])
    return (qi, (-r[0], -r[1], -r[2]))


def matrix_from_quat(q):
    """geometry.h:779-800 — rows of the 3x3, scaled by 2/|q|^2."""
    length2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]
    il = 2 / length2
    qq1 = il * q[1] * q[1]
    qq2 = il * q[2] * q[2]
    qq3 = il * q[3] * q[3]
    return ((1 - qq2 - qq3, il * (q[1] * q[2] - q[0] * q[3]), il * (q[1] * q[3] + q[0] * q[2])),
            (il * (q[1] * q[2] + q[0] * q[3]), 1 - qq1 - qq3, il * (q[2] * q[3] - q[0] * q[1])),
            (il * (q[1] * q[3] - q[0] * q[2]), il * (q[2] * q[3] + q[0] * q[1]), 1 - qq1 - qq2))


def quat_from_matrix(m):
    """geometry.h:702-743; m is 3 rows of 3."""
    tr = m[0][0] + m[1][1] + m[2][2]
    rot = [0.0, 0.0, 0.0, 0.0]
    if tr >= 0:
        rot[0] = tr + 1
        rot[1] = m[2][1] - m[1][2]
        rot[2] = m[0][2] - m[2][0]
        rot[3] = m[1][0] - m[0][1]
    elif m[1][1] > m[0][0]:
        if m[2][2] > m[1][1]:
            rot[3] = (m[2][2] - (m[0][0] + m[1][1])) + 1
            rot[1] = m[2][0] + m[0][2]
            rot[2] = m[1][2] + m[2][1]
            rot[0] = m[1][0] - m[0][1]
        else:
            rot[2] = (m[1][1] - (m[2][2] + m[0][0])) + 1
            rot[3] = m[1][2] + m[2][1]
            rot[1] = m[0][1] + m[1][0]
            rot[0] = m[0][2] - m[2][0]
    elif m[2][2] > m[0][0]:
        rot[3] = (m[2][2] - (m[0][0] + m[1][1])) + 1
        rot[1] = m[2][0] + m[0][2]
        rot[2] = m[1][2] + m[2][1]
        rot[0] = m[1][0] - m[0][1]
    else:
        rot[1] = (m[0][0] - (m[1][1] + m[2][2])) + 1
        rot[2] = m[0][1] + m[1][0]
        rot[3] = m[2][0] + m[0][2]
        rot[0] = m[2][1] - m[1][2]
    return normalize4(tuple(rot))


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def dot3(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def quat_rotate_direction(src, dst):
    """geometry.h:1257-1280 — minimal rotation taking `src` onto `dst`."""
    rd = cross(src, dst)
    fsin = math.sqrt(dot3(rd, rd))
    fcos = dot3(src, dst)
    if fsin > 0:
        inv = 1 / fsin
        return quat_from_axis_angle((rd[0] * inv, rd[1] * inv, rd[2] * inv), math.atan2(fsin, fcos))
    if fcos < 0:
        rd = (1.0, 0.0, 0.0)
        d = dot3(src, rd)
        rd = (rd[0] - src[0] * d, rd[1] - src[1] * d, rd[2] - src[2] * d)
        if dot3(rd, rd) < 1e-8:
            rd = (0.0, 0.0, 1.0)
            d = dot3(src, rd)
            rd = (rd[0] - src[0] * d, rd[1] - src[1] * d, rd[2] - src[2] * d)
        rd = normalize3(rd)
        return quat_from_axis_angle(rd, math.atan2(fsin, fcos))
    return IDENT_Q


def t_to_array8(t):
    """Flatten to the 8-double wire layout used by the C ABI: qw qx qy qz tx ty tz pad."""
    return [t[0][0], t[0][1], t[0][2], t[0][3], t[1][0], t[1][1], t[1][2], 0.0]
