// oracle.cpp — CPU restatement (IEEE double) of the reference's collision-checked-configuration path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the product
// (openrave_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library, and only as the checker / the CPU baseline.
//
// Parity pin: this restatement is checked (tests/test_oracle_vs_ref.py) against the UNMODIFIED vendored PQP
// compiled from /root/reference/plugins/pqprave/pqp by oracle/Makefile into oracle/_ref/ (identical contact
// lists and identical NumBVTests/NumTriTests on random poses), and against the reference's own known-answer
// tests: lab1 verdicts and pqp distances (tests/test_oracle_golden.py <- test/test_collision.py:183-251), the four pqp
// minDistance values of test/test_robot.py:625-660 (first one to 17 digits), and test/test_collision.py's selfcollision /
// attachedbodies / joinxml / callbacks / multiplecontacts cases (tests/test_reference_collision_tests.py).
//
// What is restated, with the reference lines it follows:
//   Transform algebra            include/openrave/geometry.h:255-262 (normalize4), :430-446 (rotate),
//                                :466-480 (operator*), :500-509 (inverse), :673-682 (quatFromAxisAngle),
//                                :779-800 (matrixFromQuat)
//   forward kinematics           src/libopenrave/kinbody.cpp:2149-2224 (limit clamp), :2286-2481 (joint loop,
//                                composition at :2478)
//   pose hand-off to PQP         plugins/pqprave/collisionPQP.h:140-147
//   BVH build                    plugins/pqprave/pqp/Build.cpp:276-330 (vertex covariance), :344-386 (split),
//                                :391-469 (recursion), :473-551 (parent-relative frames);
//                                pqp/BV.cpp:67-300 (OBB + RSS fit); pqp/MatVec.h:786-870 (Jacobi eigen);
//                                pqp/BV.h:75-85 (GetSize from the RSS fields)
//   OBB overlap                  pqp/OBB_Disjoint.h:60-211
//   triangle contact             pqp/PQP.cpp:350-459
//   dual-tree descent            pqp/PQP.cpp:482-592 (CollideRecurse), :595-652 (PQP_Collide)
//   env check order              collisionPQP.h:172-181, 314-329, 432-498, 553-566, 710-735
//   self check order             collisionPQP.h:364-386, 209-221
//
// Build variants (oracle/Makefile):
//   liboracle.so            everything above restated here
//   _ref/liboracle_ref.so   same driver, but model build / collide / distance / tolerance call the
//                           reference's own PQP objects (-DORACLE_USE_REF_PQP)
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <limits>
#include <string>
#include <vector>

#include "../include/cudacollision.h"   // POD cc_joint + option bits only

#ifdef ORACLE_USE_REF_PQP
#include <array>                    // (geometry.h uses std::array without including it)
#include <openrave/geometry.h>      // the reference's own TriTriCollision (header-only, include/openrave/geometry.h:1817-1912)
#include "PQP.h"
#endif

namespace orc {

// ------------------------------------------------------------------------------------------------
// SE(3) in the reference's quaternion form
// ------------------------------------------------------------------------------------------------
struct Xf {
    double q[4];   // w x y z
    double t[3];
};

static inline Xf xf_identity() { Xf r = {{1, 0, 0, 0}, {0, 0, 0}}; return r; }

static inline void quat_normalize(double q[4])
{
    double f = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double eps = std::numeric_limits<double>::epsilon();
    if (f < 1.0 - eps || f > 1.0 + eps) {
        f = sqrt(f);
        q[0] /= f; q[1] /= f; q[2] /= f; q[3] /= f;
    }
}

static inline void quat_rotate(const double q[4], const double r[3], double out[3])
{
    double xx = 2 * q[1] * q[1], xy = 2 * q[1] * q[2], xz = 2 * q[1] * q[3], xw = 2 * q[1] * q[0];
    double yy = 2 * q[2] * q[2], yz = 2 * q[2] * q[3], yw = 2 * q[2] * q[0];
    double zz = 2 * q[3] * q[3], zw = 2 * q[3] * q[0];
    double x = (1 - yy - zz) * r[0] + (xy - zw) * r[1] + (xz + yw) * r[2];
    double y = (xy + zw) * r[0] + (1 - xx - zz) * r[1] + (yz - xw) * r[2];
    double z = (xz - yw) * r[0] + (yz + xw) * r[1] + (1 - xx - yy) * r[2];
    out[0] = x; out[1] = y; out[2] = z;
}

static inline Xf xf_mul(const Xf& a, const Xf& b)
{
    Xf r;
    double rt[3];
    quat_rotate(a.q, b.t, rt);
    r.t[0] = a.t[0] + rt[0]; r.t[1] = a.t[1] + rt[1]; r.t[2] = a.t[2] + rt[2];
    r.q[0] = a.q[0] * b.q[0] - a.q[1] * b.q[1] - a.q[2] * b.q[2] - a.q[3] * b.q[3];
    r.q[1] = a.q[0] * b.q[1] + a.q[1] * b.q[0] + a.q[2] * b.q[3] - a.q[3] * b.q[2];
    r.q[2] = a.q[0] * b.q[2] + a.q[2] * b.q[0] + a.q[3] * b.q[1] - a.q[1] * b.q[3];
    r.q[3] = a.q[0] * b.q[3] + a.q[3] * b.q[0] + a.q[1] * b.q[2] - a.q[2] * b.q[1];
    quat_normalize(r.q);
    return r;
}

static inline void quat_from_axis_angle(const double axis[3], double angle, double q[4])
{
    double len = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    if (len == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    angle *= 0.5;
    double sang = sin(angle) / len;
    q[0] = cos(angle); q[1] = axis[0] * sang; q[2] = axis[1] * sang; q[3] = axis[2] * sang;
}

// TransformMatrix(Transform) -> PQP's R[3][3], T[3]
static inline void xf_to_RT(const Xf& x, double R[3][3], double T[3])
{
    const double* q = x.q;
    double len2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    double il = 2 / len2;
    double qq1 = il * q[1] * q[1], qq2 = il * q[2] * q[2], qq3 = il * q[3] * q[3];
    R[0][0] = 1 - qq2 - qq3;
    R[0][1] = il * (q[1] * q[2] - q[0] * q[3]);
    R[0][2] = il * (q[1] * q[3] + q[0] * q[2]);
    R[1][0] = il * (q[1] * q[2] + q[0] * q[3]);
    R[1][1] = 1 - qq1 - qq3;
    R[1][2] = il * (q[2] * q[3] - q[0] * q[1]);
    R[2][0] = il * (q[1] * q[3] - q[0] * q[2]);
    R[2][1] = il * (q[2] * q[3] + q[0] * q[1]);
    R[2][2] = 1 - qq1 - qq2;
    T[0] = x.t[0]; T[1] = x.t[1]; T[2] = x.t[2];
}

// ------------------------------------------------------------------------------------------------
// 3x3 helpers with PQP's summation order (k = 0,1,2 left to right)
// ------------------------------------------------------------------------------------------------
typedef double Mat[3][3];

static inline void mat_mul(Mat r, const Mat a, const Mat b)        // r = a b
{
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
}
static inline void mat_tmul(Mat r, const Mat a, const Mat b)       // r = a^T b
{
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            r[i][j] = a[0][i] * b[0][j] + a[1][i] * b[1][j] + a[2][i] * b[2][j];
}
static inline void mat_vec(double r[3], const Mat a, const double v[3])      // r = a v
{
    double x = a[0][0] * v[0] + a[0][1] * v[1] + a[0][2] * v[2];
    double y = a[1][0] * v[0] + a[1][1] * v[1] + a[1][2] * v[2];
    double z = a[2][0] * v[0] + a[2][1] * v[1] + a[2][2] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static inline void mat_vec_add(double r[3], const Mat a, const double v[3], const double p[3])   // r = a v + p
{
    double x = a[0][0] * v[0] + a[0][1] * v[1] + a[0][2] * v[2] + p[0];
    double y = a[1][0] * v[0] + a[1][1] * v[1] + a[1][2] * v[2] + p[1];
    double z = a[2][0] * v[0] + a[2][1] * v[1] + a[2][2] * v[2] + p[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static inline void mat_tvec(double r[3], const Mat a, const double v[3])     // r = a^T v
{
    double x = a[0][0] * v[0] + a[1][0] * v[1] + a[2][0] * v[2];
    double y = a[0][1] * v[0] + a[1][1] * v[1] + a[2][1] * v[2];
    double z = a[0][2] * v[0] + a[1][2] * v[1] + a[2][2] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static inline double dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross(double r[3], const double a[3], const double b[3])
{
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}

// ------------------------------------------------------------------------------------------------
// Triangle contact: 17-axis separating-axis test on p1-shifted coordinates; separation is strict
// (touching counts as contact); a zero axis never separates.
// ------------------------------------------------------------------------------------------------
static inline bool axis_overlaps(const double ax[3], const double p[3][3], const double q[3][3])
{
    double P0 = dot(ax, p[0]), P1 = dot(ax, p[1]), P2 = dot(ax, p[2]);
    double Q0 = dot(ax, q[0]), Q1 = dot(ax, q[1]), Q2 = dot(ax, q[2]);
    double mx1 = P0; if (P1 > mx1) mx1 = P1; if (P2 > mx1) mx1 = P2;
    double mn1 = P0; if (P1 < mn1) mn1 = P1; if (P2 < mn1) mn1 = P2;
    double mx2 = Q0; if (Q1 > mx2) mx2 = Q1; if (Q2 > mx2) mx2 = Q2;
    double mn2 = Q0; if (Q1 < mn2) mn2 = Q1; if (Q2 < mn2) mn2 = Q2;
    if (mn1 > mx2) return false;
    if (mn2 > mx1) return false;
    return true;
}

// If `min_overlap` is non-NULL and the triangles are in contact it receives the smallest normalised
// overlap over the non-degenerate axes (the oracle's penetration measure).
static bool tri_contact(const double P[3][3], const double Q[3][3], double* min_overlap = NULL)
{
    double p[3][3], q[3][3], e[3][3], f[3][3], n1[3], m1[3];
    for (int k = 0; k < 3; ++k) {
        p[0][k] = P[0][k] - P[0][k]; p[1][k] = P[1][k] - P[0][k]; p[2][k] = P[2][k] - P[0][k];
        q[0][k] = Q[0][k] - P[0][k]; q[1][k] = Q[1][k] - P[0][k]; q[2][k] = Q[2][k] - P[0][k];
    }
    for (int k = 0; k < 3; ++k) {
        e[0][k] = p[1][k] - p[0][k]; e[1][k] = p[2][k] - p[1][k]; e[2][k] = p[0][k] - p[2][k];
        f[0][k] = q[1][k] - q[0][k]; f[1][k] = q[2][k] - q[1][k]; f[2][k] = q[0][k] - q[2][k];
    }
    cross(n1, e[0], e[1]);
    cross(m1, f[0], f[1]);
    double axes[17][3];
    memcpy(axes[0], n1, sizeof(n1));
    memcpy(axes[1], m1, sizeof(m1));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cross(axes[2 + 3 * i + j], e[i], f[j]);
    for (int i = 0; i < 3; ++i) cross(axes[11 + i], e[i], n1);
    for (int i = 0; i < 3; ++i) cross(axes[14 + i], f[i], m1);
    for (int a = 0; a < 17; ++a)
        if (!axis_overlaps(axes[a], p, q)) return false;
    if (min_overlap) {
        double best = 1e300;
        for (int a = 0; a < 17; ++a) {
            double len = sqrt(dot(axes[a], axes[a]));
            if (len < 1e-14) continue;
            double P0 = dot(axes[a], p[0]), P1 = dot(axes[a], p[1]), P2 = dot(axes[a], p[2]);
            double Q0 = dot(axes[a], q[0]), Q1 = dot(axes[a], q[1]), Q2 = dot(axes[a], q[2]);
            double mx1 = fmax(P0, fmax(P1, P2)), mn1 = fmin(P0, fmin(P1, P2));
            double mx2 = fmax(Q0, fmax(Q1, Q2)), mn2 = fmin(Q0, fmin(Q1, Q2));
            double ov = fmin(mx1 - mn2, mx2 - mn1) / len;
            if (ov < best) best = ov;
        }
        *min_overlap = best;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Exact triangle-triangle distance (used only by the contact-band classifier of the restated build;
// the _ref build calls PQP_Tolerance).  Standard decomposition: min over the 9 edge-edge segment
// distances and the 6 vertex-face distances; 0 if tri_contact.
// ------------------------------------------------------------------------------------------------
static double seg_seg_dist2(const double p1[3], const double q1[3], const double p2[3], const double q2[3])
{
    double d1[3] = {q1[0] - p1[0], q1[1] - p1[1], q1[2] - p1[2]};
    double d2[3] = {q2[0] - p2[0], q2[1] - p2[1], q2[2] - p2[2]};
    double r[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    double s, t;
    const double EPS = 1e-300;
    if (a <= EPS && e <= EPS) { return dot(r, r); }
    if (a <= EPS) { s = 0; t = f / e; t = fmin(fmax(t, 0.0), 1.0); }
    else {
        double c = dot(d1, r);
        if (e <= EPS) { t = 0; s = fmin(fmax(-c / a, 0.0), 1.0); }
        else {
            double b = dot(d1, d2);
            double denom = a * e - b * b;
            if (denom > 0) s = fmin(fmax((b * f - c * e) / denom, 0.0), 1.0); else s = 0;
            t = (b * s + f) / e;
            if (t < 0) { t = 0; s = fmin(fmax(-c / a, 0.0), 1.0); }
            else if (t > 1) { t = 1; s = fmin(fmax((b - c) / a, 0.0), 1.0); }
        }
    }
    double c1[3] = {p1[0] + d1[0] * s, p1[1] + d1[1] * s, p1[2] + d1[2] * s};
    double c2[3] = {p2[0] + d2[0] * t, p2[1] + d2[1] * t, p2[2] + d2[2] * t};
    double d[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
    return dot(d, d);
}

static double point_tri_dist2(const double p[3], const double T[3][3])
{
    // Ericson, Real-Time Collision Detection 5.1.5 (closest point on triangle)
    const double* a = T[0]; const double* b = T[1]; const double* c = T[2];
    double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
    double ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    double ap[3] = {p[0] - a[0], p[1] - a[1], p[2] - a[2]};
    double cp[3];
    double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) { memcpy(cp, a, sizeof(cp)); }
    else {
        double bp[3] = {p[0] - b[0], p[1] - b[1], p[2] - b[2]};
        double d3 = dot(ab, bp), d4 = dot(ac, bp);
        if (d3 >= 0 && d4 <= d3) { memcpy(cp, b, sizeof(cp)); }
        else {
            double vc = d1 * d4 - d3 * d2;
            if (vc <= 0 && d1 >= 0 && d3 <= 0) {
                double v = d1 / (d1 - d3);
                for (int k = 0; k < 3; ++k) cp[k] = a[k] + v * ab[k];
            } else {
                double cpv[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
                double d5 = dot(ab, cpv), d6 = dot(ac, cpv);
                if (d6 >= 0 && d5 <= d6) { memcpy(cp, c, sizeof(cp)); }
                else {
                    double vb = d5 * d2 - d1 * d6;
                    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
                        double w = d2 / (d2 - d6);
                        for (int k = 0; k < 3; ++k) cp[k] = a[k] + w * ac[k];
                    } else {
                        double va = d3 * d6 - d5 * d4;
                        if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
                            double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
                            for (int k = 0; k < 3; ++k) cp[k] = b[k] + w * (c[k] - b[k]);
                        } else {
                            double denom = 1.0 / (va + vb + vc);
                            double v = vb * denom, w = vc * denom;
                            for (int k = 0; k < 3; ++k) cp[k] = a[k] + ab[k] * v + ac[k] * w;
                        }
                    }
                }
            }
        }
    }
    double d[3] = {p[0] - cp[0], p[1] - cp[1], p[2] - cp[2]};
    return dot(d, d);
}

static double tri_tri_dist(const double P[3][3], const double Q[3][3])
{
    if (tri_contact(P, Q)) return 0.0;
    double best = 1e300;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double d = seg_seg_dist2(P[i], P[(i + 1) % 3], Q[j], Q[(j + 1) % 3]);
            if (d < best) best = d;
        }
    for (int i = 0; i < 3; ++i) {
        double d = point_tri_dist2(P[i], Q); if (d < best) best = d;
        d = point_tri_dist2(Q[i], P); if (d < best) best = d;
    }
    return sqrt(best);
}

// ------------------------------------------------------------------------------------------------
// Restated PQP model: OBB tree, one triangle per leaf, parent-relative node frames
// ------------------------------------------------------------------------------------------------
struct Tri { double p[3][3]; int id; };

struct Node {
    Mat R;             // orientation (parent-relative after build)
    double To[3];      // OBB centre (parent-relative after build)
    double d[3];       // OBB half sizes
    double size;       // sqrt(l0^2 + l1^2) + 2 r from the RSS fit: the descent rule compares this
    int first_child;   // >= 0: index of first child; < 0: -(tri index + 1)
    bool leaf() const { return first_child < 0; }
};

struct CollideStats { long long bv_tests, tri_tests, contacts; };

static void jacobi_eigen(Mat vout, double dout[3], Mat a)
{
    const int n = 3;
    double b[3], z[3], d[3];
    Mat v = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int ip = 0; ip < n; ip++) { b[ip] = a[ip][ip]; d[ip] = a[ip][ip]; z[ip] = 0.0; }
    for (int i = 0; i < 50; i++) {
        double sm = 0.0;
        for (int ip = 0; ip < n; ip++) for (int iq = ip + 1; iq < n; iq++) sm += fabs(a[ip][iq]);
        if (sm == 0.0) { memcpy(vout, v, sizeof(Mat)); memcpy(dout, d, sizeof(d)); return; }
        double tresh = (i < 3) ? 0.2 * sm / (n * n) : 0.0;
        for (int ip = 0; ip < n; ip++) for (int iq = ip + 1; iq < n; iq++) {
            double g = 100.0 * fabs(a[ip][iq]);
            if ((i > 3) && (fabs(d[ip]) + g == fabs(d[ip])) && (fabs(d[iq]) + g == fabs(d[iq]))) {
                a[ip][iq] = 0.0;
            } else if (fabs(a[ip][iq]) > tresh) {
                double h = d[iq] - d[ip], t;
                if (fabs(h) + g == fabs(h)) t = (a[ip][iq]) / h;
                else {
                    double theta = 0.5 * h / (a[ip][iq]);
                    t = 1.0 / (fabs(theta) + sqrt(1.0 + theta * theta));
                    if (theta < 0.0) t = -t;
                }
                double c = 1.0 / sqrt(1 + t * t);
                double s = t * c;
                double tau = s / (1.0 + c);
                h = t * a[ip][iq];
                z[ip] -= h; z[iq] += h; d[ip] -= h; d[iq] += h;
                a[ip][iq] = 0.0;
#define ORC_ROT(m, i0, j0, k0, l0) { double gg = m[i0][j0], hh = m[k0][l0]; m[i0][j0] = gg - s * (hh + gg * tau); m[k0][l0] = hh + s * (gg - hh * tau); }
                for (int j = 0; j < ip; j++) ORC_ROT(a, j, ip, j, iq)
                for (int j = ip + 1; j < iq; j++) ORC_ROT(a, ip, j, j, iq)
                for (int j = iq + 1; j < n; j++) ORC_ROT(a, ip, j, iq, j)
                for (int j = 0; j < n; j++) ORC_ROT(v, j, ip, j, iq)
#undef ORC_ROT
            }
        }
        for (int ip = 0; ip < n; ip++) { b[ip] += z[ip]; d[ip] = b[ip]; z[ip] = 0.0; }
    }
}

class Model {
public:
    std::vector<Tri> tris;
    std::vector<Node> nodes;
    int num_nodes;
    Model() : num_nodes(0) {}

    void add_tri(const double* a, const double* b, const double* c, int id)
    {
        Tri t;
        for (int k = 0; k < 3; ++k) { t.p[0][k] = a[k]; t.p[1][k] = b[k]; t.p[2][k] = c[k]; }
        t.id = id;
        tris.push_back(t);
    }

    void build()
    {
        nodes.assign(2 * tris.size() - 1, Node());
        num_nodes = 1;
        build_node(0, 0, (int)tris.size());
        Mat I = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        double zero[3] = {0, 0, 0};
        to_parent_relative(0, I, zero);
    }

private:
    static double max2(double a, double b) { return a > b ? a : b; }

    void fit(Node& b, const Mat O, int first, int n)
    {
        memcpy(b.R, O, sizeof(Mat));
        int np = 3 * n;
        std::vector<double> Pv(3 * np);
        double (*P)[3] = (double (*)[3]) &Pv[0];
        for (int i = 0; i < n; ++i)
            for (int v = 0; v < 3; ++v) mat_tvec(P[3 * i + v], b.R, tris[first + i].p[v]);
        // OBB: centre/half sizes of the projected extent
        double minx, maxx, miny, maxy, minz, maxz, c[3];
        minx = maxx = P[0][0]; miny = maxy = P[0][1]; minz = maxz = P[0][2];
        for (int i = 1; i < np; i++) {
            if (P[i][0] < minx) minx = P[i][0]; else if (P[i][0] > maxx) maxx = P[i][0];
            if (P[i][1] < miny) miny = P[i][1]; else if (P[i][1] > maxy) maxy = P[i][1];
            if (P[i][2] < minz) minz = P[i][2]; else if (P[i][2] > maxz) maxz = P[i][2];
        }
        c[0] = 0.5 * (maxx + minx); c[1] = 0.5 * (maxy + miny); c[2] = 0.5 * (maxz + minz);
        mat_vec(b.To, b.R, c);
        b.d[0] = 0.5 * (maxx - minx); b.d[1] = 0.5 * (maxy - miny); b.d[2] = 0.5 * (maxz - minz);

        // RSS: only its side lengths and radius are needed here (they drive the descent rule)
        minz = maxz = P[0][2];
        for (int i = 1; i < np; i++) { if (P[i][2] < minz) minz = P[i][2]; else if (P[i][2] > maxz) maxz = P[i][2]; }
        double r = 0.5 * (maxz - minz), radsqr = r * r, cz = 0.5 * (maxz + minz);
        int minindex = 0, maxindex = 0;
        for (int i = 1; i < np; i++) { if (P[i][0] < P[minindex][0]) minindex = i; else if (P[i][0] > P[maxindex][0]) maxindex = i; }
        double dz = P[minindex][2] - cz;
        minx = P[minindex][0] + sqrt(max2(radsqr - dz * dz, 0));
        dz = P[maxindex][2] - cz;
        maxx = P[maxindex][0] - sqrt(max2(radsqr - dz * dz, 0));
        for (int i = 0; i < np; i++) if (P[i][0] < minx) { dz = P[i][2] - cz; double x = P[i][0] + sqrt(max2(radsqr - dz * dz, 0)); if (x < minx) minx = x; }
        for (int i = 0; i < np; i++) if (P[i][0] > maxx) { dz = P[i][2] - cz; double x = P[i][0] - sqrt(max2(radsqr - dz * dz, 0)); if (x > maxx) maxx = x; }
        minindex = maxindex = 0;
        for (int i = 1; i < np; i++) { if (P[i][1] < P[minindex][1]) minindex = i; else if (P[i][1] > P[maxindex][1]) maxindex = i; }
        dz = P[minindex][2] - cz;
        miny = P[minindex][1] + sqrt(max2(radsqr - dz * dz, 0));
        dz = P[maxindex][2] - cz;
        maxy = P[maxindex][1] - sqrt(max2(radsqr - dz * dz, 0));
        for (int i = 0; i < np; i++) if (P[i][1] < miny) { dz = P[i][2] - cz; double y = P[i][1] + sqrt(max2(radsqr - dz * dz, 0)); if (y < miny) miny = y; }
        for (int i = 0; i < np; i++) if (P[i][1] > maxy) { dz = P[i][2] - cz; double y = P[i][1] - sqrt(max2(radsqr - dz * dz, 0)); if (y > maxy) maxy = y; }
        // uncovered corners
        const double a = sqrt(0.5);
        for (int i = 0; i < np; i++) {
            double dx, dy, u, t;
            if (P[i][0] > maxx) {
                if (P[i][1] > maxy) {
                    dx = P[i][0] - maxx; dy = P[i][1] - maxy; u = dx * a + dy * a;
                    t = (a * u - dx) * (a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
                    u = u - sqrt(max2(radsqr - t, 0));
                    if (u > 0) { maxx += u * a; maxy += u * a; }
                } else if (P[i][1] < miny) {
                    dx = P[i][0] - maxx; dy = P[i][1] - miny; u = dx * a - dy * a;
                    t = (a * u - dx) * (a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
                    u = u - sqrt(max2(radsqr - t, 0));
                    if (u > 0) { maxx += u * a; miny -= u * a; }
                }
            } else if (P[i][0] < minx) {
                if (P[i][1] > maxy) {
                    dx = P[i][0] - minx; dy = P[i][1] - maxy; u = dy * a - dx * a;
                    t = (-a * u - dx) * (-a * u - dx) + (a * u - dy) * (a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
                    u = u - sqrt(max2(radsqr - t, 0));
                    if (u > 0) { minx -= u * a; maxy += u * a; }
                } else if (P[i][1] < miny) {
                    dx = P[i][0] - minx; dy = P[i][1] - miny; u = -dx * a - dy * a;
                    t = (-a * u - dx) * (-a * u - dx) + (-a * u - dy) * (-a * u - dy) + (cz - P[i][2]) * (cz - P[i][2]);
                    u = u - sqrt(max2(radsqr - t, 0));
                    if (u > 0) { minx -= u * a; miny -= u * a; }
                }
            }
        }
        double l0 = maxx - minx; if (l0 < 0) l0 = 0;
        double l1 = maxy - miny; if (l1 < 0) l1 = 0;
        b.size = sqrt(l0 * l0 + l1 * l1) + 2 * r;
    }

    int partition(int first, int n, const double axis[3], double coord)
    {
        int c1 = 0;
        for (int i = 0; i < n; i++) {
            Tri& t = tris[first + i];
            double p[3];
            for (int k = 0; k < 3; ++k) { p[k] = t.p[0][k]; p[k] = p[k] + t.p[1][k]; p[k] = p[k] + t.p[2][k]; }
            double x = dot(p, axis);
            x /= 3.0;
            if (x <= coord) { Tri tmp = t; t = tris[first + c1]; tris[first + c1] = tmp; c1++; }
        }
        if (c1 == 0 || c1 == n) c1 = n / 2;
        return c1;
    }

    void build_node(int bn, int first, int n)
    {
        // covariance of the triangle vertices (not area weighted)
        double S1[3] = {0, 0, 0};
        Mat S2 = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int i = 0; i < n; i++) {
            const Tri& t = tris[first + i];
            const double* p1 = t.p[0]; const double* p2 = t.p[1]; const double* p3 = t.p[2];
            S1[0] += p1[0] + p2[0] + p3[0];
            S1[1] += p1[1] + p2[1] + p3[1];
            S1[2] += p1[2] + p2[2] + p3[2];
            S2[0][0] += (p1[0] * p1[0] + p2[0] * p2[0] + p3[0] * p3[0]);
            S2[1][1] += (p1[1] * p1[1] + p2[1] * p2[1] + p3[1] * p3[1]);
            S2[2][2] += (p1[2] * p1[2] + p2[2] * p2[2] + p3[2] * p3[2]);
            S2[0][1] += (p1[0] * p1[1] + p2[0] * p2[1] + p3[0] * p3[1]);
            S2[0][2] += (p1[0] * p1[2] + p2[0] * p2[2] + p3[0] * p3[2]);
            S2[1][2] += (p1[1] * p1[2] + p2[1] * p2[2] + p3[1] * p3[2]);
        }
        double nn = (double)(3 * n);
        Mat C;
        C[0][0] = S2[0][0] - S1[0] * S1[0] / nn;
        C[1][1] = S2[1][1] - S1[1] * S1[1] / nn;
        C[2][2] = S2[2][2] - S1[2] * S1[2] / nn;
        C[0][1] = S2[0][1] - S1[0] * S1[1] / nn;
        C[1][2] = S2[1][2] - S1[1] * S1[2] / nn;
        C[0][2] = S2[0][2] - S1[0] * S1[2] / nn;
        C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
        Mat E; double s[3];
        jacobi_eigen(E, s, C);
        int mn, md, mx;
        if (s[0] > s[1]) { mx = 0; mn = 1; } else { mn = 0; mx = 1; }
        if (s[2] < s[mn]) { md = mn; mn = 2; } else if (s[2] > s[mx]) { md = mx; mx = 2; } else { md = 2; }
        Mat R;
        R[0][0] = E[0][mx]; R[1][0] = E[1][mx]; R[2][0] = E[2][mx];
        R[0][1] = E[0][md]; R[1][1] = E[1][md]; R[2][1] = E[2][md];
        R[0][2] = E[1][mx] * E[2][md] - E[1][md] * E[2][mx];
        R[1][2] = E[0][md] * E[2][mx] - E[0][mx] * E[2][md];
        R[2][2] = E[0][mx] * E[1][md] - E[0][md] * E[1][mx];
        fit(nodes[bn], R, first, n);
        if (n == 1) {
            nodes[bn].first_child = -(first + 1);
        } else {
            nodes[bn].first_child = num_nodes;
            num_nodes += 2;
            double axis[3] = {R[0][0], R[1][0], R[2][0]};
            double mean[3] = {0, 0, 0};
            for (int i = 0; i < n; i++) {
                const Tri& t = tris[first + i];
                mean[0] += t.p[0][0] + t.p[1][0] + t.p[2][0];
                mean[1] += t.p[0][1] + t.p[1][1] + t.p[2][1];
                mean[2] += t.p[0][2] + t.p[1][2] + t.p[2][2];
            }
            mean[0] /= nn; mean[1] /= nn; mean[2] /= nn;
            double coord = dot(axis, mean);
            int nfirst = partition(first, n, axis, coord);
            int fc = nodes[bn].first_child;
            build_node(fc, first, nfirst);
            build_node(fc + 1, first + nfirst, n - nfirst);
        }
    }

    void to_parent_relative(int bn, const Mat parentR, const double parentTo[3])
    {
        if (!nodes[bn].leaf()) {
            int fc = nodes[bn].first_child;
            // children first: they need this node's world-relative frame
            Mat Rw; memcpy(Rw, nodes[bn].R, sizeof(Mat));
            double Tw[3] = {nodes[bn].To[0], nodes[bn].To[1], nodes[bn].To[2]};
            to_parent_relative(fc, Rw, Tw);
            to_parent_relative(fc + 1, Rw, Tw);
        }
        Mat Rpc; double Tpc[3];
        mat_tmul(Rpc, parentR, nodes[bn].R);
        memcpy(nodes[bn].R, Rpc, sizeof(Mat));
        Tpc[0] = nodes[bn].To[0] - parentTo[0]; Tpc[1] = nodes[bn].To[1] - parentTo[1]; Tpc[2] = nodes[bn].To[2] - parentTo[2];
        mat_tvec(nodes[bn].To, parentR, Tpc);
    }
};

// 15-axis OBB separating-axis test; box A = identity frame with half sizes a, box B = (B,T) with half
// sizes b; |B| entries fattened by 1e-6; `<=` comparisons (touching boxes overlap).
static bool obb_overlap(const Mat B, const double T[3], const double a[3], const double b[3])
{
    const double reps = 1e-6;
    Mat Bf;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Bf[i][j] = fabs(B[i][j]); Bf[i][j] += reps; }
    double t, s;
    t = fabs(T[0]);
    if (!(t <= (a[0] + b[0] * Bf[0][0] + b[1] * Bf[0][1] + b[2] * Bf[0][2]))) return false;
    s = T[0] * B[0][0] + T[1] * B[1][0] + T[2] * B[2][0]; t = fabs(s);
    if (!(t <= (b[0] + a[0] * Bf[0][0] + a[1] * Bf[1][0] + a[2] * Bf[2][0]))) return false;
    t = fabs(T[1]);
    if (!(t <= (a[1] + b[0] * Bf[1][0] + b[1] * Bf[1][1] + b[2] * Bf[1][2]))) return false;
    t = fabs(T[2]);
    if (!(t <= (a[2] + b[0] * Bf[2][0] + b[1] * Bf[2][1] + b[2] * Bf[2][2]))) return false;
    s = T[0] * B[0][1] + T[1] * B[1][1] + T[2] * B[2][1]; t = fabs(s);
    if (!(t <= (b[1] + a[0] * Bf[0][1] + a[1] * Bf[1][1] + a[2] * Bf[2][1]))) return false;
    s = T[0] * B[0][2] + T[1] * B[1][2] + T[2] * B[2][2]; t = fabs(s);
    if (!(t <= (b[2] + a[0] * Bf[0][2] + a[1] * Bf[1][2] + a[2] * Bf[2][2]))) return false;
    s = T[2] * B[1][0] - T[1] * B[2][0]; t = fabs(s);
    if (!(t <= (a[1] * Bf[2][0] + a[2] * Bf[1][0] + b[1] * Bf[0][2] + b[2] * Bf[0][1]))) return false;
    s = T[2] * B[1][1] - T[1] * B[2][1]; t = fabs(s);
    if (!(t <= (a[1] * Bf[2][1] + a[2] * Bf[1][1] + b[0] * Bf[0][2] + b[2] * Bf[0][0]))) return false;
    s = T[2] * B[1][2] - T[1] * B[2][2]; t = fabs(s);
    if (!(t <= (a[1] * Bf[2][2] + a[2] * Bf[1][2] + b[0] * Bf[0][1] + b[1] * Bf[0][0]))) return false;
    s = T[0] * B[2][0] - T[2] * B[0][0]; t = fabs(s);
    if (!(t <= (a[0] * Bf[2][0] + a[2] * Bf[0][0] + b[1] * Bf[1][2] + b[2] * Bf[1][1]))) return false;
    s = T[0] * B[2][1] - T[2] * B[0][1]; t = fabs(s);
    if (!(t <= (a[0] * Bf[2][1] + a[2] * Bf[0][1] + b[0] * Bf[1][2] + b[2] * Bf[1][0]))) return false;
    s = T[0] * B[2][2] - T[2] * B[0][2]; t = fabs(s);
    if (!(t <= (a[0] * Bf[2][2] + a[2] * Bf[0][2] + b[0] * Bf[1][1] + b[1] * Bf[1][0]))) return false;
    s = T[1] * B[0][0] - T[0] * B[1][0]; t = fabs(s);
    if (!(t <= (a[0] * Bf[1][0] + a[1] * Bf[0][0] + b[1] * Bf[2][2] + b[2] * Bf[2][1]))) return false;
    s = T[1] * B[0][1] - T[0] * B[1][1]; t = fabs(s);
    if (!(t <= (a[0] * Bf[1][1] + a[1] * Bf[0][1] + b[0] * Bf[2][2] + b[2] * Bf[2][0]))) return false;
    s = T[1] * B[0][2] - T[0] * B[1][2]; t = fabs(s);
    if (!(t <= (a[0] * Bf[1][2] + a[1] * Bf[0][2] + b[0] * Bf[2][1] + b[1] * Bf[2][0]))) return false;
    return true;
}

struct Contact { int id1, id2; };

struct CollideCtx {
    const Model* o1; const Model* o2;
    Mat R; double T[3];            // model 2 -> model 1
    bool first_only;
    CollideStats st;
    std::vector<Contact>* out;
    double max_penetration;         // max over contacts of the min normalised SAT overlap
    bool want_penetration;
};

static void descend(CollideCtx& c, const Mat R, const double T[3], int b1, int b2)
{
    c.st.bv_tests++;
    const Node& n1 = c.o1->nodes[b1];
    const Node& n2 = c.o2->nodes[b2];
    if (!obb_overlap(R, T, n1.d, n2.d)) return;
    bool l1 = n1.leaf(), l2 = n2.leaf();
    if (l1 && l2) {
        c.st.tri_tests++;
        const Tri& t1 = c.o1->tris[-n1.first_child - 1];
        const Tri& t2 = c.o2->tris[-n2.first_child - 1];
        double q[3][3];
        for (int v = 0; v < 3; ++v) mat_vec_add(q[v], c.R, t2.p[v], c.T);
        double ov = 0;
        if (tri_contact(t1.p, q, c.want_penetration ? &ov : NULL)) {
            c.st.contacts++;
            if (c.out) { Contact k = {t1.id, t2.id}; c.out->push_back(k); }
            if (c.want_penetration && ov > c.max_penetration) c.max_penetration = ov;
        }
        return;
    }
    Mat Rc; double Tc[3], Tt[3];
    if (l2 || (!l1 && (n1.size > n2.size))) {
        int c1 = n1.first_child;
        for (int k = 0; k < 2; ++k) {
            const Node& ch = c.o1->nodes[c1 + k];
            mat_tmul(Rc, ch.R, R);
            Tt[0] = T[0] - ch.To[0]; Tt[1] = T[1] - ch.To[1]; Tt[2] = T[2] - ch.To[2];
            mat_tvec(Tc, ch.R, Tt);
            descend(c, Rc, Tc, c1 + k, b2);
            if (c.first_only && c.st.contacts > 0) return;
        }
    } else {
        int c1 = n2.first_child;
        for (int k = 0; k < 2; ++k) {
            const Node& ch = c.o2->nodes[c1 + k];
            mat_mul(Rc, R, ch.R);
            mat_vec_add(Tc, R, ch.To, T);
            descend(c, Rc, Tc, b1, c1 + k);
            if (c.first_only && c.st.contacts > 0) return;
        }
    }
}

// PQP_Collide: relative pose, root frames, recursive descent.  Returns the number of contacts.
static long long collide_models(const Mat R1, const double T1[3], const Model& m1, const Mat R2, const double T2[3],
                                const Model& m2, bool first_only, CollideStats* st, std::vector<Contact>* out,
                                double* max_pen)
{
    CollideCtx c;
    c.o1 = &m1; c.o2 = &m2; c.first_only = first_only; c.out = out;
    c.st.bv_tests = c.st.tri_tests = c.st.contacts = 0;
    c.want_penetration = max_pen != NULL; c.max_penetration = 0;
    mat_tmul(c.R, R1, R2);
    double Tt[3] = {T2[0] - T1[0], T2[1] - T1[1], T2[2] - T1[2]};
    mat_tvec(c.T, R1, Tt);
    Mat Rtemp, R; double T[3];
    mat_mul(Rtemp, c.R, m2.nodes[0].R);
    mat_tmul(R, m1.nodes[0].R, Rtemp);
    mat_vec_add(Tt, c.R, m2.nodes[0].To, c.T);
    Tt[0] -= m1.nodes[0].To[0]; Tt[1] -= m1.nodes[0].To[1]; Tt[2] -= m1.nodes[0].To[2];
    mat_tvec(T, m1.nodes[0].R, Tt);
    descend(c, R, T, 0, 0);
    if (st) *st = c.st;
    if (max_pen) *max_pen = c.max_penetration;
    return c.st.contacts;
}

// "closer than tol?" by an inflated dual-tree descent + exact triangle distance (restated build's
// replacement for PQP_Tolerance; only used by the contact-band classifier).
static bool within_tolerance(const Model& m1, const Model& m2, const Mat Rrel, const double Trel[3],
                             const Mat R, const double T[3], int b1, int b2, double tol)
{
    const Node& n1 = m1.nodes[b1];
    const Node& n2 = m2.nodes[b2];
    double a[3] = {n1.d[0] + tol, n1.d[1] + tol, n1.d[2] + tol};
    if (!obb_overlap(R, T, a, n2.d)) return false;
    bool l1 = n1.leaf(), l2 = n2.leaf();
    if (l1 && l2) {
        const Tri& t1 = m1.tris[-n1.first_child - 1];
        const Tri& t2 = m2.tris[-n2.first_child - 1];
        double q[3][3];
        for (int v = 0; v < 3; ++v) mat_vec_add(q[v], Rrel, t2.p[v], Trel);
        return tri_tri_dist(t1.p, q) <= tol;
    }
    Mat Rc; double Tc[3], Tt[3];
    if (l2 || (!l1 && (n1.size > n2.size))) {
        int c1 = n1.first_child;
        for (int k = 0; k < 2; ++k) {
            const Node& ch = m1.nodes[c1 + k];
            mat_tmul(Rc, ch.R, R);
            Tt[0] = T[0] - ch.To[0]; Tt[1] = T[1] - ch.To[1]; Tt[2] = T[2] - ch.To[2];
            mat_tvec(Tc, ch.R, Tt);
            if (within_tolerance(m1, m2, Rrel, Trel, Rc, Tc, c1 + k, b2, tol)) return true;
        }
    } else {
        int c1 = n2.first_child;
        for (int k = 0; k < 2; ++k) {
            const Node& ch = m2.nodes[c1 + k];
            mat_mul(Rc, R, ch.R);
            mat_vec_add(Tc, R, ch.To, T);
            if (within_tolerance(m1, m2, Rrel, Trel, Rc, Tc, b1, c1 + k, tol)) return true;
        }
    }
    return false;
}

// ------------------------------------------------------------------------------------------------
// Link model wrapper: restated Model, or the reference's PQP_Model in the _ref build
// ------------------------------------------------------------------------------------------------
struct LinkModel {
    bool valid;
    int ntris;
#ifdef ORACLE_USE_REF_PQP
    PQP_Model* pm;
    LinkModel() : valid(false), ntris(0), pm(NULL) {}
    ~LinkModel() { delete pm; }
#else
    Model m;
    LinkModel() : valid(false), ntris(0) {}
#endif
    void build(const double* verts, const int32_t* tri, int ntri)
    {
        ntris = ntri;
        valid = ntri > 0;
        if (!valid) return;
#ifdef ORACLE_USE_REF_PQP
        delete pm;
        pm = new PQP_Model();
        pm->BeginModel(ntri);
        for (int j = 0; j < ntri; ++j) {
            PQP_REAL p1[3], p2[3], p3[3];
            for (int k = 0; k < 3; ++k) {
                p1[k] = verts[3 * tri[3 * j] + k]; p2[k] = verts[3 * tri[3 * j + 1] + k]; p3[k] = verts[3 * tri[3 * j + 2] + k];
            }
            pm->AddTri(p1, p2, p3, j);
        }
        pm->EndModel();
#else
        m = Model();
        m.tris.reserve(ntri);
        for (int j = 0; j < ntri; ++j)
            m.add_tri(&verts[3 * tri[3 * j]], &verts[3 * tri[3 * j + 1]], &verts[3 * tri[3 * j + 2]], j);
        m.build();
#endif
    }
private:
    LinkModel(const LinkModel&);
    LinkModel& operator=(const LinkModel&);
};

struct PairResult {
    long long contacts, bv_tests, tri_tests;
    double max_penetration;
};

static PairResult collide_links(const Xf& x1, const LinkModel& l1, const Xf& x2, const LinkModel& l2, bool first_only,
                                bool want_pen, std::vector<Contact>* out = NULL)
{
    PairResult r = {0, 0, 0, 0.0};
    double R1[3][3], T1[3], R2[3][3], T2[3];
    xf_to_RT(x1, R1, T1);
    xf_to_RT(x2, R2, T2);
#ifdef ORACLE_USE_REF_PQP
    PQP_CollideResult res;
    PQP_Collide(&res, R1, T1, l1.pm, R2, T2, l2.pm, first_only ? PQP_FIRST_CONTACT : PQP_ALL_CONTACTS);
    r.contacts = res.NumPairs(); r.bv_tests = res.NumBVTests(); r.tri_tests = res.NumTriTests();
    if (out) for (int i = 0; i < res.NumPairs(); ++i) { Contact k = {res.Id1(i), res.Id2(i)}; out->push_back(k); }
    if (want_pen && res.NumPairs() > 0) {
        // penetration measure evaluated on the reference's own contact list
        double best = 0;
        std::vector<const ::Tri*> by1(l1.pm->num_tris), by2(l2.pm->num_tris);
        for (int i = 0; i < l1.pm->num_tris; ++i) by1[l1.pm->tris[i].id] = &l1.pm->tris[i];
        for (int i = 0; i < l2.pm->num_tris; ++i) by2[l2.pm->tris[i].id] = &l2.pm->tris[i];
        for (int i = 0; i < res.NumPairs(); ++i) {
            const ::Tri* a = by1[res.Id1(i)]; const ::Tri* b = by2[res.Id2(i)];
            double P[3][3], Q[3][3];
            for (int k = 0; k < 3; ++k) { P[0][k] = a->p1[k]; P[1][k] = a->p2[k]; P[2][k] = a->p3[k]; }
            mat_vec_add(Q[0], res.R, b->p1, res.T); mat_vec_add(Q[1], res.R, b->p2, res.T); mat_vec_add(Q[2], res.R, b->p3, res.T);
            double ov = 0;
            if (tri_contact(P, Q, &ov) && ov > best) best = ov;
        }
        r.max_penetration = best;
    }
#else
    CollideStats st;
    double pen = 0;
    r.contacts = collide_models(R1, T1, l1.m, R2, T2, l2.m, first_only, &st, out, want_pen ? &pen : NULL);
    r.bv_tests = st.bv_tests; r.tri_tests = st.tri_tests; r.max_penetration = pen;
#endif
    return r;
}

static bool links_within(const Xf& x1, const LinkModel& l1, const Xf& x2, const LinkModel& l2, double tol)
{
    double R1[3][3], T1[3], R2[3][3], T2[3];
    xf_to_RT(x1, R1, T1);
    xf_to_RT(x2, R2, T2);
#ifdef ORACLE_USE_REF_PQP
    PQP_ToleranceResult res;
    PQP_Tolerance(&res, R1, T1, l1.pm, R2, T2, l2.pm, tol);
    return res.CloserThanTolerance() != 0;
#else
    Mat Rrel; double Trel[3], Tt[3] = {T2[0] - T1[0], T2[1] - T1[1], T2[2] - T1[2]};
    mat_tmul(Rrel, R1, R2);
    mat_tvec(Trel, R1, Tt);
    Mat Rtemp, R; double T[3];
    mat_mul(Rtemp, Rrel, l2.m.nodes[0].R);
    mat_tmul(R, l1.m.nodes[0].R, Rtemp);
    mat_vec_add(Tt, Rrel, l2.m.nodes[0].To, Trel);
    Tt[0] -= l1.m.nodes[0].To[0]; Tt[1] -= l1.m.nodes[0].To[1]; Tt[2] -= l1.m.nodes[0].To[2];
    mat_tvec(T, l1.m.nodes[0].R, Tt);
    return within_tolerance(l1.m, l2.m, Rrel, Trel, R, T, 0, 0, tol);
#endif
}

// Exact minimum distance of two link models (PQP_Distance with rel_err = abs_err = 0, pqp/PQP.cpp:663-945): in the restated
// build a branch and bound over the two OBB trees — the separation of two boxes along A's and along B's face axes bounds
// their distance from below, the nearer child is descended first (DistanceRecurse's rule), leaves run the exact triangle
// distance above.  `best` carries the running minimum in and out.
#ifndef ORACLE_USE_REF_PQP
static double obb_gap(const Mat R, const double T[3], const double a[3], const double b[3])
{
    double ga = 0, gb = 0;
    for (int i = 0; i < 3; ++i) {
        double g = fabs(T[i]) - (a[i] + b[0] * fabs(R[i][0]) + b[1] * fabs(R[i][1]) + b[2] * fabs(R[i][2]));
        if (g > 0) ga += g * g;
        double tb = T[0] * R[0][i] + T[1] * R[1][i] + T[2] * R[2][i];
        g = fabs(tb) - (b[i] + a[0] * fabs(R[0][i]) + a[1] * fabs(R[1][i]) + a[2] * fabs(R[2][i]));
        if (g > 0) gb += g * g;
    }
    return sqrt(ga > gb ? ga : gb);
}

static void distance_descend(const Model& m1, const Model& m2, const Mat Rrel, const double Trel[3],
                             const Mat R, const double T[3], int b1, int b2, double& best)
{
    const Node& n1 = m1.nodes[b1];
    const Node& n2 = m2.nodes[b2];
    bool l1 = n1.leaf(), l2 = n2.leaf();
    if (l1 && l2) {
        const Tri& t1 = m1.tris[-n1.first_child - 1];
        const Tri& t2 = m2.tris[-n2.first_child - 1];
        double q[3][3];
        for (int v = 0; v < 3; ++v) mat_vec_add(q[v], Rrel, t2.p[v], Trel);
        double d = tri_tri_dist(t1.p, q);
        if (d < best) best = d;
        return;
    }
    Mat Rc[2]; double Tc[2][3], Tt[3], lb[2];
    int ia[2], ib[2];
    if (l2 || (!l1 && (n1.size > n2.size))) {
        int c1 = n1.first_child;
        for (int k = 0; k < 2; ++k) {
            const Node& ch = m1.nodes[c1 + k];
            mat_tmul(Rc[k], ch.R, R);
            Tt[0] = T[0] - ch.To[0]; Tt[1] = T[1] - ch.To[1]; Tt[2] = T[2] - ch.To[2];
            mat_tvec(Tc[k], ch.R, Tt);
            ia[k] = c1 + k; ib[k] = b2;
            lb[k] = obb_gap(Rc[k], Tc[k], ch.d, n2.d);
        }
    } else {
        int c1 = n2.first_child;
        for (int k = 0; k < 2; ++k) {
            const Node& ch = m2.nodes[c1 + k];
            mat_mul(Rc[k], R, ch.R);
            mat_vec_add(Tc[k], R, ch.To, T);
            ia[k] = b1; ib[k] = c1 + k;
            lb[k] = obb_gap(Rc[k], Tc[k], n1.d, ch.d);
        }
    }
    const int first = lb[1] < lb[0] ? 1 : 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int k = pass == 0 ? first : 1 - first;
        if (lb[k] < best) distance_descend(m1, m2, Rrel, Trel, Rc[k], Tc[k], ia[k], ib[k], best);
    }
}
#endif

static double links_distance(const Xf& x1, const LinkModel& l1, const Xf& x2, const LinkModel& l2, double best)
{
    double R1[3][3], T1[3], R2[3][3], T2[3];
    xf_to_RT(x1, R1, T1);
    xf_to_RT(x2, R2, T2);
#ifdef ORACLE_USE_REF_PQP
    PQP_DistanceResult res;
    PQP_Distance(&res, R1, T1, l1.pm, R2, T2, l2.pm, 0.0, 0.0);
    return res.Distance() < best ? res.Distance() : best;
#else
    Mat Rrel; double Trel[3], Tt[3] = {T2[0] - T1[0], T2[1] - T1[1], T2[2] - T1[2]};
    mat_tmul(Rrel, R1, R2);
    mat_tvec(Trel, R1, Tt);
    Mat Rtemp, R; double T[3];
    mat_mul(Rtemp, Rrel, l2.m.nodes[0].R);
    mat_tmul(R, l1.m.nodes[0].R, Rtemp);
    mat_vec_add(Tt, Rrel, l2.m.nodes[0].To, Trel);
    Tt[0] -= l1.m.nodes[0].To[0]; Tt[1] -= l1.m.nodes[0].To[1]; Tt[2] -= l1.m.nodes[0].To[2];
    mat_tvec(T, l1.m.nodes[0].R, Tt);
    if (obb_gap(R, T, l1.m.nodes[0].d, l2.m.nodes[0].d) < best) distance_descend(l1.m, l2.m, Rrel, Trel, R, T, 0, 0, best);
    return best;
#endif
}

// ------------------------------------------------------------------------------------------------
// Scene: bodies in environment order, links with models and current transforms, kinematics
// ------------------------------------------------------------------------------------------------
struct Body {
    std::string name;
    bool is_robot;
    bool removed;
    std::vector<LinkModel*> links;
    std::vector<uint8_t> enabled;
    std::vector<Xf> xf;
    std::vector<int> attached;        // body ids rigidly attached
    int ndof;
    std::vector<cc_joint> joints;
    std::vector<uint8_t> affects;     // [ndof][nlinks]
    std::vector<int32_t> nonadjacent; // packed, reference order
    std::vector<double> dofvals;      // current values of all DOFs (used for the inactive ones)
    struct Grab {                     // KinBody::_vGrabbedBodies entry (kinbodygrab.cpp)
        int body, link;
        std::vector<Xf> rel;          // pose of every grabbed link relative to the grabbing link
        std::vector<int32_t> noncolliding;   // (body, link) pairs: _listNonCollidingLinksWhenGrabbed
    };
    std::vector<Grab> grabs;
    // affine active DOFs (RobotBase::SetActiveDOFs(indices, affine, axis)): bits of OpenRAVE::DOFAffine, the rotation axis of
    // DOF_RotationAxis and _vRotationQuatLimitStart; their values follow the joint values in every row of a batch
    int affine_mask;
    double affine_axis[3];
    double affine_qstart[4];
    Body() : is_robot(false), removed(false), ndof(0), affine_mask(0) { affine_axis[0] = affine_axis[1] = 0; affine_axis[2] = 1; affine_qstart[0] = 1; affine_qstart[1] = affine_qstart[2] = affine_qstart[3] = 0; }
    ~Body() { for (size_t i = 0; i < links.size(); ++i) delete links[i]; }
};

struct Scene {
    std::vector<Body*> bodies;
    ~Scene() { for (size_t i = 0; i < bodies.size(); ++i) delete bodies[i]; }
};

static int affine_count(int m) { return ((m & 1) ? 1 : 0) + ((m & 2) ? 1 : 0) + ((m & 4) ? 1 : 0) + ((m & 8) ? 1 : 0) + ((m & 16) ? 3 : 0) + ((m & 32) ? 4 : 0); }

// RaveGetTransformFromAffineDOFValues (libopenrave.cpp:1868-1924) + the _vRotationQuatLimitStart product of
// RobotBase::SetActiveDOFValues (robot.cpp:1203-1207), applied to the body's current pose t
static Xf affine_transform(const Body& b, Xf t, const double* a)
{
    const int m = b.affine_mask;
    if (m & 1) t.t[0] = *a++;
    if (m & 2) t.t[1] = *a++;
    if (m & 4) t.t[2] = *a++;
    if (m & 8) {
        const double angle = *a++, fsin = sin(angle * 0.5);
        t.q[0] = cos(angle * 0.5); t.q[1] = b.affine_axis[0] * fsin; t.q[2] = b.affine_axis[1] * fsin; t.q[3] = b.affine_axis[2] * fsin;
    } else if (m & 16) {
        const double x = a[0], y = a[1], z = a[2], fang = sqrt(x * x + y * y + z * z);
        if (fang > 0) { const double f = sin(0.5 * fang) / fang; t.q[0] = cos(0.5 * fang); t.q[1] = f * x; t.q[2] = f * y; t.q[3] = f * z; }
        else { t.q[0] = 1; t.q[1] = t.q[2] = t.q[3] = 0; }
    } else if (m & 32) {
        double r[4] = {a[0], a[1], a[2], a[3]};
        double f = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3];
        if (f < 1 - 2.220446049250313e-16 || f > 1 + 2.220446049250313e-16) { f = sqrt(f); for (int k = 0; k < 4; ++k) r[k] /= f; }      // normalize4
        const double* s = b.affine_qstart;                                                                                                 // quatMultiply(start, r)
        t.q[0] = s[0] * r[0] - s[1] * r[1] - s[2] * r[2] - s[3] * r[3];
        t.q[1] = s[0] * r[1] + s[1] * r[0] + s[2] * r[3] - s[3] * r[2];
        t.q[2] = s[0] * r[2] + s[2] * r[0] + s[3] * r[1] - s[1] * r[3];
        t.q[3] = s[0] * r[3] + s[3] * r[0] + s[1] * r[2] - s[2] * r[1];
    }
    return t;
}

// KinBody::SetDOFValues' joint loop.  `vals` holds ALL DOF values; they are clamped here like
// kinbody.cpp:2149-2224 does (non-circular DOFs only).
static void forward_kinematics(const Body& b, const double* vals_in, const Xf& base, std::vector<Xf>& out)
{
    std::vector<double> vals(vals_in, vals_in + b.ndof);
    for (size_t k = 0; k < b.joints.size(); ++k) {
        const cc_joint& j = b.joints[k];
        if (j.dof_index >= 0 && !j.is_circular) {
            if (vals[j.dof_index] < j.lower) vals[j.dof_index] = j.lower;
            else if (vals[j.dof_index] > j.upper) vals[j.dof_index] = j.upper;
        }
    }
    out = b.xf;
    out[0] = base;
    std::vector<uint8_t> done(b.links.size(), 0);
    done[0] = 1;
    for (size_t k = 0; k < b.joints.size(); ++k) {
        const cc_joint& j = b.joints[k];
        Xf L, Rr;
        memcpy(L.q, j.left, 4 * sizeof(double)); memcpy(L.t, j.left + 4, 3 * sizeof(double));
        memcpy(Rr.q, j.right, 4 * sizeof(double)); memcpy(Rr.t, j.right + 4, 3 * sizeof(double));
        const Xf& parent = out[j.parent_link];
        if (j.is_static) {
            out[j.child_link] = xf_mul(parent, L);
            done[j.child_link] = 1;
            continue;
        }
        double v;
        if (j.mimic_dof >= 0) {
            double ev = j.mimic_a * vals[j.mimic_dof] + j.mimic_b;
            v = ev;
            if (!j.is_circular) {
                // g_fEpsilon = 1e-15 for dReal = double (openrave.h:130); g_fEpsilonJointLimit = RavePow(g_fEpsilon, 0.8),
                // g_fEpsilonEvalJointLimit = RavePow(g_fEpsilon, 0.65) (libopenrave.h:141-142); rules of kinbody.cpp:2326-2375
                static const double eps_jl = pow(1e-15, 0.8), eps_eval = pow(1e-15, 0.65);
                if (ev < j.lower) { if (ev >= j.lower - eps_jl) v = j.lower; else if (ev < j.lower - eps_eval) v = j.lower; }
                else if (ev > j.upper) { if (ev <= j.upper + eps_jl) v = j.upper; else if (ev > j.upper + eps_eval) v = j.upper; }
            }
        } else if (j.dof_index >= 0) {
            v = vals[j.dof_index];
        } else {
            v = 0.0;
            if (!j.is_circular) { if (v < j.lower) v = j.lower; else if (v > j.upper) v = j.upper; }
        }
        if (done[j.child_link]) continue;
        Xf J = xf_identity();
        if (j.type == CC_JOINT_REVOLUTE) quat_from_axis_angle(j.axis, v, J.q);
        else { J.t[0] = j.axis[0] * v; J.t[1] = j.axis[1] * v; J.t[2] = j.axis[2] * v; }
        Xf local = xf_mul(xf_mul(L, J), Rr);
        out[j.child_link] = xf_mul(parent, local);
        done[j.child_link] = 1;
    }
}

static bool is_attached(const Body& a, int other)
{
    for (size_t i = 0; i < a.attached.size(); ++i) if (a.attached[i] == other) return true;
    return false;
}

// _IsActiveLink (collisionPQP.h:710-735): with CO_ActiveDOFs only links moved by an active DOF count
static void active_links(const Body& b, const int32_t* active, int nactive, int options, std::vector<uint8_t>& out)
{
    out.assign(b.links.size(), 1);
    if (!(options & CC_CO_ActiveDOFs) || !b.is_robot) return;
    if (b.affine_mask) return;                   // "enable everything" (collisionPQP.h:716-719)
    out.assign(b.links.size(), 0);
    for (size_t l = 0; l < b.links.size(); ++l)
        for (int a = 0; a < nactive; ++a)
            if (b.affects[(size_t)active[a] * b.links.size() + l]) { out[l] = 1; break; }
}

// IsDOFInChain for the AO_ActiveDOFs pair filter (robot.cpp:2056-2066, kinbody.cpp:2607-2612)
static void self_pairs(const Body& b, const int32_t* active, int nactive, int options, std::vector<int32_t>& out)
{
    out.clear();
    size_t nl = b.links.size();
    for (size_t i = 0; i < b.nonadjacent.size(); ++i) {
        int p = b.nonadjacent[i];
        int l0 = p & 0xffff, l1 = (p >> 16) & 0xffff;
        if ((options & CC_CO_ActiveDOFs) && b.is_robot) {
            bool any = false;
            for (int a = 0; a < nactive && !any; ++a)
                any = (b.affects[(size_t)active[a] * nl + l0] == 0) != (b.affects[(size_t)active[a] * nl + l1] == 0);
            if (!any) continue;
        }
        if (!b.enabled[l0] || !b.enabled[l1]) continue;   // AO_Enabled
        out.push_back(p);
    }
    // filtered subsequences of a CompareNonAdjacentFarthest-sorted list stay sorted
}

}  // namespace orc

using namespace orc;

// ------------------------------------------------------------------------------------------------
// C ABI (mirrors the product's scene-building calls so the same flattening code feeds both)
// ------------------------------------------------------------------------------------------------
extern "C" {

void* orc_scene_create() { return new Scene(); }
void orc_scene_destroy(void* s) { delete (Scene*)s; }
int orc_uses_reference_pqp()
{
#ifdef ORACLE_USE_REF_PQP
    return 1;
#else
    return 0;
#endif
}

int orc_scene_add_body(void* sp, const char* name, int nlinks, int is_robot)
{
    Scene* s = (Scene*)sp;
    Body* b = new Body();
    b->name = name ? name : "";
    b->is_robot = is_robot != 0;
    b->links.resize(nlinks);
    for (int i = 0; i < nlinks; ++i) b->links[i] = new LinkModel();
    b->enabled.assign(nlinks, 1);
    b->xf.assign(nlinks, xf_identity());
    s->bodies.push_back(b);
    int id = (int)s->bodies.size() - 1;
    b->attached.push_back(id);
    return id;
}

int orc_body_set_link_mesh(void* sp, int body, int link, const double* verts, int nverts, const int32_t* tris, int ntris)
{
    Scene* s = (Scene*)sp; (void)nverts;
    s->bodies[body]->links[link]->build(verts, tris, ntris);
    return 0;
}

int orc_body_set_link_enabled(void* sp, int body, const uint8_t* en)
{
    Body* b = ((Scene*)sp)->bodies[body];
    b->enabled.assign(en, en + b->links.size());
    return 0;
}

int orc_body_set_link_transforms(void* sp, int body, const double* T8)
{
    Body* b = ((Scene*)sp)->bodies[body];
    for (size_t i = 0; i < b->links.size(); ++i) {
        memcpy(b->xf[i].q, T8 + 8 * i, 4 * sizeof(double));
        memcpy(b->xf[i].t, T8 + 8 * i + 4, 3 * sizeof(double));
    }
    return 0;
}

int orc_body_set_attached(void* sp, int a, int b, int attached)
{
    Scene* s = (Scene*)sp;
    for (int pass = 0; pass < 2; ++pass) {
        Body* x = s->bodies[pass ? b : a]; int o = pass ? a : b;
        bool has = is_attached(*x, o);
        if (attached && !has) x->attached.push_back(o);
        if (!attached && has) { for (size_t i = 0; i < x->attached.size(); ++i) if (x->attached[i] == o) { x->attached.erase(x->attached.begin() + i); break; } }
    }
    return 0;
}

int orc_body_set_grabbed(void* sp, int body, int grabbed, int link, const double* rel8, const int32_t* noncolliding, int n)
{
    Scene* s = (Scene*)sp;
    Body* b = s->bodies[body];
    Body::Grab* g = NULL;
    for (size_t i = 0; i < b->grabs.size(); ++i) if (b->grabs[i].body == grabbed) g = &b->grabs[i];
    if (!g) { b->grabs.push_back(Body::Grab()); g = &b->grabs.back(); }
    g->body = grabbed; g->link = link;
    const size_t ng = s->bodies[grabbed]->links.size();
    g->rel.resize(ng);
    for (size_t k = 0; k < ng; ++k) { memcpy(g->rel[k].q, rel8 + 8 * k, 32); memcpy(g->rel[k].t, rel8 + 8 * k + 4, 24); }
    g->noncolliding.assign(noncolliding, noncolliding + 2 * (size_t)n);
    return orc_body_set_attached(sp, body, grabbed, 1);
}

int orc_body_release_grabbed(void* sp, int body, int grabbed)
{
    Body* b = ((Scene*)sp)->bodies[body];
    for (size_t i = 0; i < b->grabs.size();) {
        if (grabbed < 0 || b->grabs[i].body == grabbed) { orc_body_set_attached(sp, body, b->grabs[i].body, 0); b->grabs.erase(b->grabs.begin() + i); }
        else ++i;
    }
    return 0;
}

int orc_body_set_removed(void* sp, int body, int removed) { ((Scene*)sp)->bodies[body]->removed = removed != 0; return 0; }

int orc_body_set_kinematics(void* sp, int body, int ndof, int njoints, const cc_joint* joints, const uint8_t* affects,
                            const double* current_dof_values)
{
    Body* b = ((Scene*)sp)->bodies[body];
    b->ndof = ndof;
    b->joints.assign(joints, joints + njoints);
    b->affects.assign(affects, affects + (size_t)ndof * b->links.size());
    b->dofvals.assign(current_dof_values, current_dof_values + ndof);
    return 0;
}

int orc_body_set_affine(void* sp, int body, int mask, const double* axis3, const double* qstart4)
{
    Body* b = ((Scene*)sp)->bodies[body];
    b->affine_mask = mask;
    if (axis3) for (int k = 0; k < 3; ++k) b->affine_axis[k] = axis3[k];
    if (qstart4) for (int k = 0; k < 4; ++k) b->affine_qstart[k] = qstart4[k];
    return 0;
}

int orc_body_set_nonadjacent_pairs(void* sp, int body, const int32_t* packed, int n)
{
    Body* b = ((Scene*)sp)->bodies[body];
    b->nonadjacent.assign(packed, packed + n);
    return 0;
}

// CheckCollision(link,link) x n at explicit transforms of `body` (GetNonAdjacentLinks' initial-pose probe,
// kinbody.cpp:5563-5577).  pairs = packed i|(j<<16).
int orc_check_link_pairs(void* sp, int body, const double* T8, const int32_t* packed, int n, uint8_t* hit)
{
    Body* b = ((Scene*)sp)->bodies[body];
    std::vector<Xf> xf(b->links.size());
    for (size_t i = 0; i < xf.size(); ++i) { memcpy(xf[i].q, T8 + 8 * i, 32); memcpy(xf[i].t, T8 + 8 * i + 4, 24); }
    for (int k = 0; k < n; ++k) {
        int l0 = packed[k] & 0xffff, l1 = (packed[k] >> 16) & 0xffff;
        hit[k] = 0;
        if (!b->enabled[l0] || !b->enabled[l1]) continue;
        if (!b->links[l0]->valid || !b->links[l1]->valid) continue;
        hit[k] = collide_links(xf[l0], *b->links[l0], xf[l1], *b->links[l1], false, false).contacts > 0;
    }
    return 0;
}

// Per-configuration result of the batched oracle.
//   verdict        bit0 env, bit1 self
//   first_env[3]   {query link, other body, other link} of the first colliding env pair in pqprave order, -1s if none
//   first_self[2]  {link i, link j} of the first colliding non-adjacent pair, -1s if none
//   counters[2]    NumBVTests, NumTriTests summed over the link pairs visited
//   margin         with want_band: free configs -> 1 if some visited pair is closer than `band`, else 0;
//                  colliding configs -> max penetration measure of the first colliding pair(s)
struct OrcOptions {
    int32_t options;          // cc_collision_options
    int32_t flags;            // CC_BATCH_ENV | CC_BATCH_SELF
    int32_t report;           // 1: env check visits every pair (report != NULL semantics), 0: first hit returns
    int32_t want_band;        // classify the contact band
    double band;              // 1e-4
    int32_t pose_mode;        // 1: `q` holds [n][nlinks][8] link transforms instead of DOF values
    int32_t reserved;
};

// env_pairs_out (optional): for report mode, every colliding env pair of configuration c is appended as 4 ints
// {config, query link, other body, other link}; capacity in quads; *n_env_pairs receives the count.
int orc_check_batch(void* sp, int body, const int32_t* active, int nactive, const OrcOptions* opt,
                    const double* q, int n, uint32_t* verdict, int32_t* first_env, int32_t* first_self,
                    int64_t* counters, double* margin, uint8_t* near_contact,
                    int32_t* env_pairs_out, int64_t env_pairs_cap, int64_t* n_env_pairs, double* fk_out)
{
    Scene* s = (Scene*)sp;
    Body* b = s->bodies[body];
    const size_t nl = b->links.size();
    std::vector<uint8_t> act;
    active_links(*b, active, nactive, opt->options, act);
    std::vector<int32_t> spairs;
    self_pairs(*b, active, nactive, opt->options, spairs);
    std::vector<double> vals(b->dofvals);
    std::vector<Xf> xf;
    int64_t nout = 0;
    // query links: the body's own, then the links of its grabbed bodies in grab order (same layout as the product's
    // cc_query_get_links); qb/ql name them, gbase[g] is where grabbed body g starts
    std::vector<Body*> qb; std::vector<int> ql; std::vector<size_t> gbase(b->grabs.size(), 0);
    for (size_t i = 0; i < nl; ++i) { qb.push_back(b); ql.push_back((int)i); }
    for (size_t g = 0; g < b->grabs.size(); ++g) {
        gbase[g] = qb.size();
        Body* G = s->bodies[b->grabs[g].body];
        for (size_t k = 0; k < G->links.size(); ++k) { qb.push_back(G); ql.push_back((int)k); }
    }
    const size_t nle = qb.size();
    for (int c = 0; c < n; ++c) {
        if (opt->pose_mode) {
            xf.resize(nle);
            const double* T8 = q + (size_t)c * nle * 8;
            for (size_t i = 0; i < nle; ++i) { memcpy(xf[i].q, T8 + 8 * i, 32); memcpy(xf[i].t, T8 + 8 * i + 4, 24); }
        } else {
            const size_t ncols = (size_t)nactive + affine_count(b->affine_mask);
            for (int a = 0; a < nactive; ++a) vals[active[a]] = q[(size_t)c * ncols + a];
            forward_kinematics(*b, &vals[0], b->affine_mask ? affine_transform(*b, b->xf[0], q + (size_t)c * ncols + nactive) : b->xf[0], xf);
            // KinBody::_UpdateGrabbedBodies: grabbed links follow their grabbing link
            xf.resize(nle);
            for (size_t g = 0; g < b->grabs.size(); ++g)
                for (size_t k = 0; k < b->grabs[g].rel.size(); ++k) xf[gbase[g] + k] = xf_mul(xf[b->grabs[g].link], b->grabs[g].rel[k]);
        }
        if (fk_out) {
            for (size_t i = 0; i < nle; ++i) {
                double* o = fk_out + ((size_t)c * nle + i) * 8;
                memcpy(o, xf[i].q, 32); memcpy(o + 4, xf[i].t, 24); o[7] = 0;
            }
        }
        uint32_t v = 0;
        int64_t bv = 0, tt = 0;
        int fe[3] = {-1, -1, -1}, fs[2] = {-1, -1};
        double pen = 0; bool near = false;
        if (opt->flags & CC_BATCH_ENV) {
            bool done = false;
            // CheckCollision(body): every attached body (here: the body, then its grabbed bodies) through CheckCollisionP
            // (collisionPQP.h:314-329, 432-498)
            for (int att = -1; att < (int)b->grabs.size() && !done; ++att) {
                Body* a1 = att < 0 ? b : s->bodies[b->grabs[att].body];
                const size_t base = att < 0 ? 0 : gbase[att];
                for (size_t ob = 0; ob < s->bodies.size() && !done; ++ob) {
                    Body* o = s->bodies[ob];
                    if (o->removed) continue;
                    bool attached = is_attached(*b, (int)ob);                      // GetAttached is transitive
                    for (size_t g = 0; g < b->grabs.size() && !attached; ++g)
                        attached = b->grabs[g].body == (int)ob || is_attached(*s->bodies[b->grabs[g].body], (int)ob);
                    if (attached) continue;
                    for (size_t i = 0; i < a1->links.size() && !done; ++i) {
                        for (size_t j = 0; j < o->links.size() && !done; ++j) {
                            if (!a1->enabled[i] || !o->enabled[j]) continue;
                            if (att < 0 && !act[i]) continue;
                            if (!a1->links[i]->valid || !o->links[j]->valid) continue;
                            PairResult r = collide_links(xf[base + i], *a1->links[i], o->xf[j], *o->links[j], false, opt->want_band != 0);
                            bv += r.bv_tests; tt += r.tri_tests;
                            if (r.contacts > 0) {
                                if (!(v & CC_VERDICT_ENV)) { fe[0] = (int)(base + i); fe[1] = (int)ob; fe[2] = (int)j; }
                                v |= CC_VERDICT_ENV;
                                if (r.max_penetration > pen) pen = r.max_penetration;
                                if (env_pairs_out && nout < env_pairs_cap) {
                                    int32_t* e = env_pairs_out + 4 * nout;
                                    e[0] = c; e[1] = (int)(base + i); e[2] = (int)ob; e[3] = (int)j;
                                }
                                if (opt->report) nout++;
                                if (!opt->report) done = true;
                            } else if (opt->want_band && !near) {
                                near = links_within(xf[base + i], *a1->links[i], o->xf[j], *o->links[j], opt->band);
                            }
                        }
                    }
                }
            }
        }
        if (opt->flags & CC_BATCH_SELF) {
            // one link pair of the query (DoPQP: both enabled, both active, both with a model); true = contact
            struct SelfCheck {
                const std::vector<Body*>& qb; const std::vector<int>& ql; const std::vector<Xf>& xf; const std::vector<uint8_t>& act;
                size_t nl; const OrcOptions* opt; int64_t& bv; int64_t& tt; double& pen; bool& near; uint32_t& v; int* fs;
                bool operator()(size_t a, size_t b2) const
                {
                    const LinkModel& la = *qb[a]->links[ql[a]]; const LinkModel& lb = *qb[b2]->links[ql[b2]];
                    if (!qb[a]->enabled[ql[a]] || !qb[b2]->enabled[ql[b2]]) return false;
                    if ((a < nl && !act[a]) || (b2 < nl && !act[b2])) return false;
                    if (!la.valid || !lb.valid) return false;
                    PairResult r = collide_links(xf[a], la, xf[b2], lb, false, opt->want_band != 0);
                    bv += r.bv_tests; tt += r.tri_tests;
                    if (r.contacts > 0) {
                        v |= CC_VERDICT_SELF; fs[0] = (int)a; fs[1] = (int)b2;
                        if (r.max_penetration > pen) pen = r.max_penetration;
                        return true;
                    }
                    if (opt->want_band && !near) near = links_within(xf[a], la, xf[b2], lb, opt->band);
                    return false;
                }
            } check = {qb, ql, xf, act, nl, opt, bv, tt, pen, near, v, fs};
            bool hit = false;
            // CheckStandaloneSelfCollision (collisionPQP.h:364-386)
            for (size_t k = 0; k < spairs.size() && !hit; ++k) hit = check(spairs[k] & 0xffff, (spairs[k] >> 16) & 0xffff);
            // grabbed bodies (KinBody::CheckSelfCollision, kinbodycollision.cpp:83-205)
            for (size_t g1 = 0; g1 < b->grabs.size() && !hit; ++g1) {
                const Body::Grab& G1 = b->grabs[g1];
                Body* B1 = s->bodies[G1.body];
                bool en1 = false;
                for (size_t k = 0; k < B1->enabled.size(); ++k) en1 |= B1->enabled[k] != 0;
                if (!en1) continue;
                for (size_t k = 0; k + 1 < G1.noncolliding.size() && !hit; k += 2) {
                    // the listed link: one of the body's own, or a link of another body it still grabs
                    size_t a = nle;
                    if (G1.noncolliding[k] == body) a = (size_t)G1.noncolliding[k + 1];
                    else for (size_t g = 0; g < b->grabs.size(); ++g) if (b->grabs[g].body == G1.noncolliding[k]) a = gbase[g] + G1.noncolliding[k + 1];
                    if (a >= nle) continue;
                    for (size_t l = 0; l < B1->links.size() && !hit; ++l) hit = check(a, gbase[g1] + l);
                }
                // grabbedBody1.CheckSelfCollision: its own non-adjacent list
                for (size_t k = 0; k < B1->nonadjacent.size() && !hit; ++k)
                    hit = check(gbase[g1] + (B1->nonadjacent[k] & 0xffff), gbase[g1] + ((B1->nonadjacent[k] >> 16) & 0xffff));
                for (size_t g2 = g1 + 1; g2 < b->grabs.size() && !hit; ++g2) {
                    const Body::Grab& G2 = b->grabs[g2];
                    Body* B2 = s->bodies[G2.body];
                    bool en2 = false;
                    for (size_t k = 0; k < B2->enabled.size(); ++k) en2 |= B2->enabled[k] != 0;
                    if (!en2) continue;
                    for (size_t l2 = 0; l2 < B2->links.size() && !hit; ++l2) {
                        if (!B2->enabled[l2]) continue;
                        bool in1 = false;
                        for (size_t k = 0; k + 1 < G1.noncolliding.size(); k += 2) in1 |= G1.noncolliding[k] == G2.body && G1.noncolliding[k + 1] == (int)l2;
                        if (in1) continue;
                        for (size_t l1 = 0; l1 < B1->links.size() && !hit; ++l1) {
                            if (!B1->enabled[l1]) continue;
                            bool in2 = false;
                            for (size_t k = 0; k + 1 < G2.noncolliding.size(); k += 2) in2 |= G2.noncolliding[k] == G1.body && G2.noncolliding[k + 1] == (int)l1;
                            if (in2) hit = check(gbase[g1] + l1, gbase[g2] + l2);
                        }
                    }
                }
            }
        }
        if (verdict) verdict[c] = v;
        if (first_env) { first_env[3 * c] = fe[0]; first_env[3 * c + 1] = fe[1]; first_env[3 * c + 2] = fe[2]; }
        if (first_self) { first_self[2 * c] = fs[0]; first_self[2 * c + 1] = fs[1]; }
        if (counters) { counters[2 * c] = bv; counters[2 * c + 1] = tt; }
        if (margin) margin[c] = pen;
        if (near_contact) near_contact[c] = near ? 1 : 0;
    }
    if (n_env_pairs) *n_env_pairs = nout;
    return 0;
}

// CO_Distance over a batch (collisionPQP.h:622-648 applied to every pair the query visits): the exact minimum separation per
// configuration over the env pairs (flags & 1) and the self pairs (flags & 2), and the pair attaining it as
// {query link, other body, other link} (other body = the query body for self pairs).  Grabbed bodies follow the same
// query-link layout as orc_check_batch.
int orc_distance_batch(void* sp, int body, const int32_t* active, int nactive, const OrcOptions* opt,
                       const double* q, int n, double max_distance, double* min_distance, int32_t* min_pair3)
{
    Scene* s = (Scene*)sp;
    Body* b = s->bodies[body];
    const size_t nl = b->links.size();
    std::vector<uint8_t> act;
    active_links(*b, active, nactive, opt->options, act);
    std::vector<int32_t> spairs;
    self_pairs(*b, active, nactive, opt->options, spairs);
    std::vector<double> vals(b->dofvals);
    std::vector<Xf> xf;
    std::vector<Body*> qb; std::vector<int> ql; std::vector<size_t> gbase(b->grabs.size(), 0);
    for (size_t i = 0; i < nl; ++i) { qb.push_back(b); ql.push_back((int)i); }
    for (size_t g = 0; g < b->grabs.size(); ++g) {
        gbase[g] = qb.size();
        Body* G = s->bodies[b->grabs[g].body];
        for (size_t k = 0; k < G->links.size(); ++k) { qb.push_back(G); ql.push_back((int)k); }
    }
    const size_t nle = qb.size();
    for (int c = 0; c < n; ++c) {
        if (opt->pose_mode) {
            xf.resize(nle);
            const double* T8 = q + (size_t)c * nle * 8;
            for (size_t i = 0; i < nle; ++i) { memcpy(xf[i].q, T8 + 8 * i, 32); memcpy(xf[i].t, T8 + 8 * i + 4, 24); }
        } else {
            for (int a = 0; a < nactive; ++a) vals[active[a]] = q[(size_t)c * nactive + a];
            forward_kinematics(*b, &vals[0], b->xf[0], xf);
            xf.resize(nle);
            for (size_t g = 0; g < b->grabs.size(); ++g)
                for (size_t k = 0; k < b->grabs[g].rel.size(); ++k) xf[gbase[g] + k] = xf_mul(xf[b->grabs[g].link], b->grabs[g].rel[k]);
        }
        double best = max_distance > 0 ? max_distance : 1e30;
        int bp[3] = {-1, -1, -1};
        if (opt->flags & CC_BATCH_ENV) {
            for (int att = -1; att < (int)b->grabs.size(); ++att) {
                Body* a1 = att < 0 ? b : s->bodies[b->grabs[att].body];
                const size_t base = att < 0 ? 0 : gbase[att];
                for (size_t ob = 0; ob < s->bodies.size(); ++ob) {
                    Body* o = s->bodies[ob];
                    if (o->removed) continue;
                    bool attached = is_attached(*b, (int)ob);
                    for (size_t g = 0; g < b->grabs.size() && !attached; ++g)
                        attached = b->grabs[g].body == (int)ob || is_attached(*s->bodies[b->grabs[g].body], (int)ob);
                    if (attached) continue;
                    for (size_t i = 0; i < a1->links.size(); ++i)
                        for (size_t j = 0; j < o->links.size(); ++j) {
                            if (!a1->enabled[i] || !o->enabled[j] || (att < 0 && !act[i])) continue;
                            if (!a1->links[i]->valid || !o->links[j]->valid) continue;
                            double d = links_distance(xf[base + i], *a1->links[i], o->xf[j], *o->links[j], best);
                            if (d < best) { best = d; bp[0] = (int)(base + i); bp[1] = (int)ob; bp[2] = (int)j; }
                        }
                }
            }
        }
        if (opt->flags & CC_BATCH_SELF) {
            for (size_t k = 0; k < spairs.size(); ++k) {
                int l0 = spairs[k] & 0xffff, l1 = (spairs[k] >> 16) & 0xffff;
                if (!act[l0] || !act[l1] || !b->links[l0]->valid || !b->links[l1]->valid) continue;
                double d = links_distance(xf[l0], *b->links[l0], xf[l1], *b->links[l1], best);
                if (d < best) { best = d; bp[0] = l0; bp[1] = body; bp[2] = l1; }
            }
        }
        min_distance[c] = best;
        if (min_pair3) { min_pair3[3 * c] = bp[0]; min_pair3[3 * c + 1] = bp[1]; min_pair3[3 * c + 2] = bp[2]; }
    }
    return 0;
}

// Per-pair queries on a batch (parity of CO_UseTolerance and of the per-pair mode of the product's distance kernel): for
// every configuration and every listed (query link, other body, other link) triple, the exact distance clipped at `cap`
// (cap <= 0: unbounded) and PQP_Tolerance's verdict at `tolerance` (the reference's own routine in the _ref build,
// pqp/PQP.cpp:1000-1180; the restated inflated descent otherwise).  Query links follow the robot-only layout.
int orc_pair_queries(void* sp, int body, const int32_t* active, int nactive, const double* q, int n, const int32_t* triples,
                     int ntriples, double cap, double tolerance, double* dist_out, uint8_t* within_out)
{
    Scene* s = (Scene*)sp;
    Body* b = s->bodies[body];
    std::vector<double> vals(b->dofvals);
    std::vector<Xf> xf;
    for (int c = 0; c < n; ++c) {
        for (int a = 0; a < nactive; ++a) vals[active[a]] = q[(size_t)c * nactive + a];
        forward_kinematics(*b, &vals[0], b->xf[0], xf);
        for (int k = 0; k < ntriples; ++k) {
            const int la = triples[3 * k], ob = triples[3 * k + 1], lb = triples[3 * k + 2];
            const LinkModel& A = *b->links[la];
            const LinkModel& Bm = *s->bodies[ob]->links[lb];
            const Xf& xb = ob == body ? xf[lb] : s->bodies[ob]->xf[lb];
            double d = cap > 0 ? cap : 1e30;
            bool w = false;
            if (A.valid && Bm.valid) {
                if (dist_out) d = links_distance(xf[la], A, xb, Bm, d);
                if (within_out) w = links_within(xf[la], A, xb, Bm, tolerance);
            }
            if (dist_out) dist_out[(size_t)c * ntriples + k] = d;
            if (within_out) within_out[(size_t)c * ntriples + k] = w ? 1 : 0;
        }
    }
    return 0;
}

// Contact-band class of single link pairs (the per-pair form of orc_check_batch's want_band, for the CO_AllLinkCollisions
// parity test): out[c][k] = 0 free and farther apart than `band`, 1 free but within `band`, 2 colliding with a penetration
// measure <= `band`, 3 colliding deeper than that.
int orc_pair_band(void* sp, int body, const int32_t* active, int nactive, const double* q, int n, const int32_t* triples,
                  int ntriples, double band, uint8_t* out)
{
    Scene* s = (Scene*)sp;
    Body* b = s->bodies[body];
    std::vector<double> vals(b->dofvals);
    std::vector<Xf> xf;
    for (int c = 0; c < n; ++c) {
        for (int a = 0; a < nactive; ++a) vals[active[a]] = q[(size_t)c * nactive + a];
        forward_kinematics(*b, &vals[0], b->xf[0], xf);
        for (int k = 0; k < ntriples; ++k) {
            const int la = triples[3 * k], ob = triples[3 * k + 1], lb = triples[3 * k + 2];
            const LinkModel& A = *b->links[la];
            const LinkModel& Bm = *s->bodies[ob]->links[lb];
            const Xf& xb = ob == body ? xf[lb] : s->bodies[ob]->xf[lb];
            uint8_t cls = 0;
            if (A.valid && Bm.valid) {
                PairResult r = collide_links(xf[la], A, xb, Bm, false, true);
                if (r.contacts > 0) cls = r.max_penetration <= band ? 2 : 3;
                else cls = links_within(xf[la], A, xb, Bm, band) ? 1 : 0;
            }
            out[(size_t)c * ntriples + k] = cls;
        }
    }
    return 0;
}

// The reference's contact routine for one intersecting triangle pair (collisionPQP.h:591-603 calls it per PQP contact):
// the UNMODIFIED OpenRAVE::geometry::TriTriCollision in the _ref build; -1 in the restated build (contacts are pinned on the
// reference routine only).  u, v: 3 x xyz world coordinates.
int orc_tri_tri_collision(const double* u, const double* v, double* pos, double* norm)
{
#ifdef ORACLE_USE_REF_PQP
    using OpenRAVE::geometry::RaveVector;
    RaveVector<double> u1(u[0], u[1], u[2]), u2(u[3], u[4], u[5]), u3(u[6], u[7], u[8]);
    RaveVector<double> v1(v[0], v[1], v[2]), v2(v[3], v[4], v[5]), v3(v[6], v[7], v[8]), p, n;
    if (!OpenRAVE::geometry::TriTriCollision(u1, u2, u3, v1, v2, v3, p, n)) return 0;
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z; norm[0] = n.x; norm[1] = n.y; norm[2] = n.z;
    return 1;
#else
    (void)u; (void)v; (void)pos; (void)norm;
    return -1;
#endif
}

// Single-thread timing helper for the CPU baseline: runs orc_check_batch semantics (no band, no outputs
// except verdicts) and returns elapsed seconds.
double orc_time_batch(void* sp, int body, const int32_t* active, int nactive, int options, int flags,
                      const double* q, int n, uint32_t* verdict)
{
    OrcOptions o; memset(&o, 0, sizeof(o));
    o.options = options; o.flags = flags; o.report = 0; o.want_band = 0; o.band = 0; o.pose_mode = 0;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    orc_check_batch(sp, body, active, nactive, &o, q, n, verdict, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL, NULL);
    std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// The serial per-sample loop of DynamicsCollisionConstraint::Check (src/libopenrave/planningutils.cpp:3405-3454,
// _SetAndCheckState :2317-2368) behind the planner library's cc_group_check_fn signature: every group is walked segment
// by segment, configuration by configuration, in the given order, and abandoned at its first colliding configuration —
// exactly what the reference's planners pay for an edge.  `user` points at an OrcGroupChecker.
struct OrcGroupChecker {
    void* scene;
    const int32_t* active;
    int32_t body, nactive, options, reserved;
    uint64_t configs_checked;     // accumulated over calls
    double seconds;               // accumulated over calls
};

int orc_group_check(void* user, const double* q, int nconfigs, int ndof, const int32_t* group_begin, int ngroups,
                    const int32_t* segment_begin, int nsegments, int32_t* first_fail_segment, int32_t* fail_verdict,
                    int32_t* first_fail_config)
{
    OrcGroupChecker* u = (OrcGroupChecker*)user;
    (void)nconfigs; (void)nsegments;
    if (ndof != u->nactive) return 1;
    OrcOptions o; memset(&o, 0, sizeof(o));
    o.options = u->options; o.flags = CC_BATCH_ENV | CC_BATCH_SELF;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    for (int g = 0; g < ngroups; ++g) {
        first_fail_segment[g] = -1; fail_verdict[g] = 0; first_fail_config[g] = -1;
        for (int sgm = group_begin[g]; sgm < group_begin[g + 1] && first_fail_segment[g] < 0; ++sgm) {
            for (int c = segment_begin[sgm]; c < segment_begin[sgm + 1]; ++c) {
                uint32_t v = 0;
                orc_check_batch(u->scene, u->body, u->active, u->nactive, &o, q + (size_t)c * ndof, 1, &v, NULL, NULL, NULL, NULL, NULL,
                                NULL, 0, NULL, NULL);
                u->configs_checked++;
                if (v) { first_fail_segment[g] = sgm - group_begin[g]; fail_verdict[g] = (int32_t)v; first_fail_config[g] = c; break; }
            }
        }
    }
    u->seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return 0;
}

// Raw model-vs-model query for pinning the restatement against the reference build:
// returns contacts; fills counters {bv, tri}; optional sorted contact list as id pairs.
int64_t orc_collide_pair(void* sp, int body1, int link1, const double* T1_8, int body2, int link2, const double* T2_8,
                         int first_only, int64_t* counters, int32_t* contacts_out, int64_t cap)
{
    Scene* s = (Scene*)sp;
    Xf x1, x2;
    memcpy(x1.q, T1_8, 32); memcpy(x1.t, T1_8 + 4, 24);
    memcpy(x2.q, T2_8, 32); memcpy(x2.t, T2_8 + 4, 24);
    std::vector<Contact> out;
    PairResult r = collide_links(x1, *s->bodies[body1]->links[link1], x2, *s->bodies[body2]->links[link2], first_only != 0, false, &out);
    if (counters) { counters[0] = r.bv_tests; counters[1] = r.tri_tests; }
    if (contacts_out)
        for (int64_t i = 0; i < (int64_t)out.size() && i < cap; ++i) { contacts_out[2 * i] = out[i].id1; contacts_out[2 * i + 1] = out[i].id2; }
    return r.contacts;
}

#ifdef ORACLE_USE_REF_PQP
// PQP_Distance with the plugin's tolerances (collisionPQP.h:48-49,630) for the distance goldens.
double orc_ref_distance(void* sp, int body1, int link1, const double* T1_8, int body2, int link2, const double* T2_8,
                        double rel_err, double abs_err)
{
    Scene* s = (Scene*)sp;
    Xf x1, x2;
    memcpy(x1.q, T1_8, 32); memcpy(x1.t, T1_8 + 4, 24);
    memcpy(x2.q, T2_8, 32); memcpy(x2.t, T2_8 + 4, 24);
    double R1[3][3], T1[3], R2[3][3], T2[3];
    xf_to_RT(x1, R1, T1); xf_to_RT(x2, R2, T2);
    PQP_DistanceResult res;
    PQP_Distance(&res, R1, T1, s->bodies[body1]->links[link1]->pm, R2, T2, s->bodies[body2]->links[link2]->pm, rel_err, abs_err);
    return res.Distance();
}
#endif

int orc_link_valid(void* sp, int body, int link) { return ((Scene*)sp)->bodies[body]->links[link]->valid ? 1 : 0; }

}  // extern "C"
