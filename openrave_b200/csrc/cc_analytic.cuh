// cc_analytic.cuh — analytic primitives (fclrave semantics): box, sphere, cylinder and capsule geometries
// collide as the exact convex shapes instead of their tessellation (plugins/fclrave/fclspace.cpp:475-602 builds fcl::Box /
// Sphere / Cylinder / Capsule; fcl::collide runs GJK on them, fclcollision.cpp:77 GST_INDEP).  Off by default: pqprave, the
// primary reference, sees one triangle soup per link (Appendix B of SURVEY.md).
//
// One kernel, one THREAD per (configuration, analytic link pair): world-box broadphase, then
//   primitive x primitive   boolean GJK on the two convex shapes
//   primitive x mesh        descent of the mesh's box tree with the primitive's oriented bounding box (15-axis SAT of
//                           cc_device.cuh), boolean GJK of the primitive against each surviving triangle
// Results are merged into the outputs the traversal kernel has already written (atomicOr / atomicMin), so the first colliding
// pair in reference order and the complete masks stay what CheckCollisionP's loop order makes them.
// FCL itself is not in the reference tree (un-vendored, unpinned fork): "parity unpinned"; the tests pin this file against
// closed forms and against the tessellated answer within the sagitta of the tessellation.
#pragma once
#include "cc_device.cuh"

namespace ccdev {

struct Prim {                   // 64 bytes; pose in the link frame
    float R[9];
    float t[3];
    float dims[3];              // box: half extents; sphere: r; cylinder / capsule: r, half height (axis z)
    int32_t type;               // cc_geom_type: 1 box, 2 sphere, 3 cylinder, 4 capsule
};
static_assert(sizeof(Prim) == 64, "Prim layout");

struct AnalyticParams {
    const uint32_t* pairs;      // slot_a | slot_b << 16 in the plan's slot space
    const int32_t* orig;        // reference-order pair index
    int32_t npairs, nenv;       // env pairs first
    const int32_t* slot_prim;   // [nslots + 1] first primitive of every slot (count 0: the link collides as its mesh)
    const Prim* prims;
};

// a convex shape placed in the common (base-centred world) frame
struct Shape {
    float R[9], t[3];
    float d0, d1, d2;
    int type;                   // 1..4 as above, 0: triangle with vertices v0 = t, v1 = (R[0..2]), v2 = (R[3..5]) in the common frame
};

__device__ __forceinline__ V3 shape_support(const Shape& s, V3 d)
{
    if (s.type == 0) {
        const V3 a = v3(s.t[0], s.t[1], s.t[2]), b = v3(s.R[0], s.R[1], s.R[2]), c = v3(s.R[3], s.R[4], s.R[5]);
        const float da = dot(a, d), db = dot(b, d), dc = dot(c, d);
        return da >= db ? (da >= dc ? a : c) : (db >= dc ? b : c);
    }
    // direction in the shape's frame
    const float lx = s.R[0] * d.x + s.R[3] * d.y + s.R[6] * d.z;
    const float ly = s.R[1] * d.x + s.R[4] * d.y + s.R[7] * d.z;
    const float lz = s.R[2] * d.x + s.R[5] * d.y + s.R[8] * d.z;
    float px, py, pz;
    if (s.type == 1) { px = lx >= 0.f ? s.d0 : -s.d0; py = ly >= 0.f ? s.d1 : -s.d1; pz = lz >= 0.f ? s.d2 : -s.d2; }
    else if (s.type == 2) {
        const float n = sqrtf(lx * lx + ly * ly + lz * lz);
        const float f = n > 0.f ? s.d0 / n : 0.f;
        px = lx * f; py = ly * f; pz = lz * f;
    } else if (s.type == 3) {
        const float n = sqrtf(lx * lx + ly * ly);
        const float f = n > 0.f ? s.d0 / n : 0.f;
        px = lx * f; py = ly * f; pz = lz >= 0.f ? s.d1 : -s.d1;
    } else {
        const float n = sqrtf(lx * lx + ly * ly + lz * lz);
        const float f = n > 0.f ? s.d0 / n : 0.f;
        px = lx * f; py = ly * f; pz = lz * f + (lz >= 0.f ? s.d1 : -s.d1);
    }
    return v3(fmaf(s.R[0], px, fmaf(s.R[1], py, fmaf(s.R[2], pz, s.t[0]))), fmaf(s.R[3], px, fmaf(s.R[4], py, fmaf(s.R[5], pz, s.t[1]))),
              fmaf(s.R[6], px, fmaf(s.R[7], py, fmaf(s.R[8], pz, s.t[2]))));
}

__device__ __forceinline__ V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 triple(V3 a, V3 b, V3 c) { return cross(cross(a, b), c); }      // (a x b) x c

// Boolean GJK: do the two convex shapes share a point?  Touching counts as contact (like the triangle test, which separates
// only on a strict gap); a search that has not decided after kGjkIters iterations is reported as contact — it is then within
// rounding of touching.
constexpr int kGjkIters = 48;
__device__ bool gjk_intersect(const Shape& A, const Shape& B)
{
    V3 d = v3(B.t[0] - A.t[0], B.t[1] - A.t[1], B.t[2] - A.t[2]);
    if (A.type == 0 || B.type == 0 || dot(d, d) < 1e-20f) d = v3(1.f, 0.f, 0.f);
    V3 s0, s1, s2, s3;                                   // simplex, newest first
    s0 = shape_support(A, d) - shape_support(B, neg(d));
    int n = 1;
    d = neg(s0);
    for (int it = 0; it < kGjkIters; ++it) {
        const float dd = dot(d, d);
        if (dd < 1e-24f) return true;                    // the origin lies on the simplex
        const V3 a = shape_support(A, d) - shape_support(B, neg(d));
        if (dot(a, d) < 0.f) return false;               // a separating direction: no contact (strictly apart)
        s3 = s2; s2 = s1; s1 = s0; s0 = a; ++n;
        const V3 ao = neg(s0);
        if (n == 2) {
            const V3 ab = s1 - s0;
            if (dot(ab, ao) > 0.f) d = triple(ab, ao, ab);
            else { n = 1; d = ao; }
        } else if (n == 3) {
            const V3 ab = s1 - s0, ac = s2 - s0, abc = cross(ab, ac);
            if (dot(cross(abc, ac), ao) > 0.f) {
                if (dot(ac, ao) > 0.f) { s1 = s2; n = 2; d = triple(ac, ao, ac); }
                else if (dot(ab, ao) > 0.f) { n = 2; d = triple(ab, ao, ab); }
                else { n = 1; d = ao; }
            } else if (dot(cross(ab, abc), ao) > 0.f) {
                if (dot(ab, ao) > 0.f) { n = 2; d = triple(ab, ao, ab); }
                else { n = 1; d = ao; }
            } else if (dot(abc, ao) > 0.f) d = abc;
            else { const V3 tmp = s1; s1 = s2; s2 = tmp; d = neg(abc); }
        } else {
            const V3 ab = s1 - s0, ac = s2 - s0, ad = s3 - s0;
            const V3 abc = cross(ab, ac), acd = cross(ac, ad), adb = cross(ad, ab);
            // orient the three faces that contain the newest point away from the fourth vertex
            const bool oabc = dot(abc, ao) * (dot(abc, ad) > 0.f ? -1.f : 1.f) > 0.f;
            const bool oacd = dot(acd, ao) * (dot(acd, ab) > 0.f ? -1.f : 1.f) > 0.f;
            const bool oadb = dot(adb, ao) * (dot(adb, ac) > 0.f ? -1.f : 1.f) > 0.f;
            if (!oabc && !oacd && !oadb) return true;    // the origin is inside the tetrahedron
            // keep the face the origin is in front of (the first such face; the triangle case refines next iteration)
            if (oabc) { n = 3; d = dot(abc, ad) > 0.f ? neg(abc) : abc; }
            else if (oacd) { s1 = s2; s2 = s3; n = 3; d = dot(acd, ab) > 0.f ? neg(acd) : acd; }
            else { const V3 b = s1; s1 = s3; s2 = b; n = 3; d = dot(adb, ac) > 0.f ? neg(adb) : adb; }
        }
    }
    return true;
}

__device__ __forceinline__ void compose_rt(const float* A, const float* B, float* o)       // o = A * B, 12 floats each (R row major | t)
{
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) o[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
        o[9 + i] = fmaf(A[3 * i], B[9], fmaf(A[3 * i + 1], B[10], fmaf(A[3 * i + 2], B[11], A[9 + i])));
    }
}

__device__ __forceinline__ void load_pose(const Params& P, const float* xf, int slot, float* o)
{
    const float4* e = slot < P.ndyn ? (const float4*)(xf + 12 * slot) : (const float4*)(P.st_xf + 12 * (size_t)(slot - P.ndyn));
    const float4 e0 = e[0], e1 = e[1], e2 = e[2];
    o[0] = e0.x; o[1] = e0.y; o[2] = e0.z; o[3] = e0.w; o[4] = e1.x; o[5] = e1.y; o[6] = e1.z; o[7] = e1.w;
    o[8] = e2.x; o[9] = e2.y; o[10] = e2.z; o[11] = e2.w;
}

__device__ __forceinline__ Shape place_prim(const Prim& g, const float* link_pose)
{
    Shape s;
    float loc[12], w[12];
#pragma unroll
    for (int i = 0; i < 9; ++i) loc[i] = g.R[i];
    loc[9] = g.t[0]; loc[10] = g.t[1]; loc[11] = g.t[2];
    compose_rt(link_pose, loc, w);
#pragma unroll
    for (int i = 0; i < 9; ++i) s.R[i] = w[i];
    s.t[0] = w[9]; s.t[1] = w[10]; s.t[2] = w[11];
    s.d0 = g.dims[0]; s.d1 = g.dims[1]; s.d2 = g.dims[2];
    s.type = g.type;
    return s;
}

// half extents of the primitive's bounding box in its own frame
__device__ __forceinline__ void prim_box(const Prim& g, float& ex, float& ey, float& ez)
{
    if (g.type == 1) { ex = g.dims[0]; ey = g.dims[1]; ez = g.dims[2]; }
    else if (g.type == 2) { ex = ey = ez = g.dims[0]; }
    else if (g.type == 3) { ex = ey = g.dims[0]; ez = g.dims[1]; }
    else { ex = ey = g.dims[0]; ez = g.dims[1] + g.dims[0]; }
}

// one primitive (placed by `gpose` in the common frame) against the mesh of slot `ms` (link pose `mpose`)
__device__ bool prim_vs_mesh(const Params& P, const Prim& g, const float* gpose_link, const float* mpose, int mroot)
{
    // primitive frame -> mesh link frame: M^-1 * (L * g)
    float gl[12], gw[12], rel[12], minv[12];
#pragma unroll
    for (int i = 0; i < 9; ++i) gl[i] = g.R[i];
    gl[9] = g.t[0]; gl[10] = g.t[1]; gl[11] = g.t[2];
    compose_rt(gpose_link, gl, gw);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) minv[3 * i + j] = mpose[3 * j + i];
    for (int i = 0; i < 3; ++i) minv[9 + i] = -(minv[3 * i] * mpose[9] + minv[3 * i + 1] * mpose[10] + minv[3 * i + 2] * mpose[11]);
    compose_rt(minv, gw, rel);
    float ex, ey, ez;
    prim_box(g, ex, ey, ez);
    const float pad = 1e-5f;                             // the mesh boxes are inflated the same way
    Shape S;
#pragma unroll
    for (int i = 0; i < 9; ++i) S.R[i] = rel[i];
    S.t[0] = rel[9]; S.t[1] = rel[10]; S.t[2] = rel[11];
    S.d0 = g.dims[0]; S.d1 = g.dims[1]; S.d2 = g.dims[2]; S.type = g.type;
    int stack[64];
    int sp = 0;
    stack[sp++] = mroot;
    while (sp > 0) {
        const int nd = stack[--sp];
        const float4 c = __ldg(&P.nodes[2 * nd]), e = __ldg(&P.nodes[2 * nd + 1]);
        if (!box_overlap(rel, rel + 9, c.x, c.y, c.z, e.x, e.y, e.z, 0.f, 0.f, 0.f, ex + pad, ey + pad, ez + pad)) continue;
        const int left = __float_as_int(c.w);
        if (left < 0) {
            const float4* tv = P.tris + 3 * (size_t)(~left);
            const float4 a = __ldg(tv), b = __ldg(tv + 1), cc = __ldg(tv + 2);
            Shape T;
            T.type = 0;
            T.t[0] = a.x; T.t[1] = a.y; T.t[2] = a.z;
            T.R[0] = b.x; T.R[1] = b.y; T.R[2] = b.z; T.R[3] = cc.x; T.R[4] = cc.y; T.R[5] = cc.z;
            T.R[6] = T.R[7] = T.R[8] = 0.f; T.d0 = T.d1 = T.d2 = 0.f;
            if (gjk_intersect(S, T)) return true;
        } else if (sp <= 62) { stack[sp++] = left; stack[sp++] = left + 1; }
        else return true;                                // (deeper than any tree built here: be conservative)
    }
    return false;
}

__global__ void __launch_bounds__(128)
cc_analytic_kernel(const __grid_constant__ Params P, const __grid_constant__ AnalyticParams A)
{
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (long long)P.n * A.npairs) return;
    const int c = (int)(w / A.npairs), k = (int)(w - (long long)c * A.npairs);
    const bool is_env = k < A.nenv;
    if (is_env ? !(P.flags & kFlagEnv) : !(P.flags & kFlagSelf)) return;
    const uint32_t pr = __ldg(&A.pairs[k]);
    const int sa = pr & 0xffff, sb = pr >> 16;
    const float* xf = P.xf32 + (size_t)c * P.ndyn * 12;
    const float* ab = P.aabb + (size_t)c * P.ndyn * 8;
    {
        const float4* pa = (const float4*)(sa < P.ndyn ? ab + 8 * sa : P.st_aabb + 8 * (size_t)(sa - P.ndyn));
        const float4* pb = (const float4*)(sb < P.ndyn ? ab + 8 * sb : P.st_aabb + 8 * (size_t)(sb - P.ndyn));
        const float4 ac = pa[0], ae = pa[1], bc = pb[0], be = pb[1];
        if (!(fabsf(ac.x - bc.x) <= ae.x + be.x && fabsf(ac.y - bc.y) <= ae.y + be.y && fabsf(ac.z - bc.z) <= ae.z + be.z)) return;
    }
    float pa[12], pb[12];
    load_pose(P, xf, sa, pa);
    load_pose(P, xf, sb, pb);
    const int a0 = __ldg(&A.slot_prim[sa]), a1 = __ldg(&A.slot_prim[sa + 1]);
    const int b0 = __ldg(&A.slot_prim[sb]), b1 = __ldg(&A.slot_prim[sb + 1]);
    bool hit = false;
    if (a1 > a0 && b1 > b0) {
        for (int i = a0; i < a1 && !hit; ++i) {
            const Shape SA = place_prim(A.prims[i], pa);
            for (int j = b0; j < b1 && !hit; ++j) hit = gjk_intersect(SA, place_prim(A.prims[j], pb));
        }
    } else if (a1 > a0) {
        const int root = __ldg(&P.slot_root[sb]);
        for (int i = a0; i < a1 && !hit; ++i) hit = prim_vs_mesh(P, A.prims[i], pa, pb, root);
    } else if (b1 > b0) {
        const int root = __ldg(&P.slot_root[sa]);
        for (int j = b0; j < b1 && !hit; ++j) hit = prim_vs_mesh(P, A.prims[j], pb, pa, root);
    }
    if (!hit) return;
    const int p = __ldg(&A.orig[k]);
    if (P.verdict) atomicOr(&P.verdict[c], is_env ? 1u : 2u);
    if (!(P.flags & kFlagFirstHit)) {
        if (is_env) { if (P.first_env) atomicMin((unsigned int*)&P.first_env[c], (unsigned int)p); }
        else if (P.first_self) atomicMin((unsigned int*)&P.first_self[c], (unsigned int)p);
    }
    if (P.pair_mask && (P.flags & kFlagAllPairs)) atomicOr(&P.pair_mask[(size_t)c * P.mask_words + (p >> 5)], 1u << (p & 31));
}

}  // namespace ccdev
