// cc_device.cuh — device side of the batched collision checker (sm_100a only).
//
// One persistent kernel, `cc_check_kernel`: each WARP owns one configuration at a time (pulled from a
// global atomic counter) and runs the whole reference pipeline for it without touching HBM in between:
//
//   1. batched FK       KinBody::SetDOFValues' joint loop (kinbody.cpp:2286-2481) in FP64: lanes build the
//                       per-joint local transforms (L * J(q) * R) in parallel, then compose level by level
//                       through shared memory; link poses stay in shared memory (FP32 R|t, base-centred).
//   2. broadphase       one lane per (query link, other link) pair of the bound pair table: relative pose +
//                       15-axis SAT of the two root boxes.  Survivors get a "slot" holding their relative pose.
//   3. traversal        warp-cooperative dual-BVH descent: a LIFO of (nodeA, nodeB, slot) items lives in
//                       shared memory, every iteration pops up to 32 items (one per lane), tests them and
//                       pushes children by ballot/popc compaction.  Nodes are 32-byte link-frame AABBs fetched
//                       as two 16-byte loads; they are served from L2 (the whole scene is a few MB).
//   4. narrowphase      leaf x leaf -> PQP's 17-axis triangle SAT (pqp/PQP.cpp:350-459) in FP32 on
//                       p1-shifted coordinates in link A's frame.
//   5. verdict          first colliding pair in reference order per class (env / self) by warp min-reduction,
//                       optional complete pair bitmask.
//
// Conservative pruning: node boxes are inflated at build time and |R| entries get +1e-6 like
// OBB_Disjoint.h:67-78, so FP32 rounding can never prune a true contact; the verdict is the OR of the
// triangle tests of unpruned pairs, exactly like the reference's (BVH-shape independent).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ccdev {

constexpr int kWarpsPerBlock = 4;
constexpr int kStackCap = 512;       // items of shared-memory LIFO per warp
constexpr int kSpillChunk = 256;     // items moved to / from the global spill area at a time
constexpr int kNumSlots = 64;        // candidate pairs traversed together
constexpr unsigned kFull = 0xffffffffu;

// flags (low bits mirror cc_batch_flags)
constexpr int kFlagEnv = 1, kFlagSelf = 2, kFlagAllPairs = 4, kFlagPoseMode = 8, kFlagFkOnly = 16, kFlagCounters = 32;

struct FkStep {                 // one joint of the chain program, 256 bytes
    int32_t type;               // 1 revolute, 2 prismatic
    int32_t parent, child;
    int32_t src_col;            // column of q feeding this joint (own DOF or the mimic source); -1: constant
    int32_t flags;              // bit0 static, bit1 mimic, bit2 own-circular, bit3 source-circular
    int32_t pad[3];
    double src_const;           // value when src_col < 0
    double src_lower, src_upper;// limits of the source DOF (KinBody::SetDOFValues clamps DOF values first)
    double a, b;                // mimic: value = a*src + b
    double lower, upper;        // own limits (mimic clamp rule)
    double L[8];                // left transform  {qw qx qy qz tx ty tz 0}
    double R[8];                // right transform
    double axis[4];
    double pad2;
};

struct Params {
    // scene (device pointers)
    const float4* nodes;        // 2 x float4 per node: {cx cy cz | left} {ex ey ez | right}
    const float4* tris;         // 3 x float4 per triangle (link frame)
    const int32_t* slot_root;   // root node of every link slot
    const float* env_xf;        // [nslots - nq][12] FP32 R(9, row major)|t(3), base-centred
    // query
    const uint32_t* pairs;      // device pair list: a | (b << 16), b = slot index
    const int32_t* pair_orig;   // original (reference-order) index of each device pair
    const FkStep* steps;
    const int32_t* level_start; // [max_depth + 2] offsets into steps (steps are sorted by depth)
    const uint32_t* hoisted_mask;
    int32_t nq, nslots, npairs_dev, nenv_dev, nenv_total, nsteps, max_depth, ncols;
    int32_t hoisted_first_env, hoisted_first_self, mask_words, flags;
    double base[8];             // pose of link 0 (kept, kinbody.cpp:2281)
    double center[3];           // subtracted from every translation before the FP32 conversion
    float tri_eps;              // conservative margin of the triangle test (metres), 0 = reference's strict test
    // io
    const double* q;            // [n][ncols]   (or [n][nq][8] in pose mode)
    int32_t n;
    uint32_t* verdict;
    int32_t* first_env;
    int32_t* first_self;
    uint32_t* pair_mask;
    double* fk_out;             // [n][nq][8] or NULL
    // runtime
    int32_t* work_counter;
    unsigned long long* stats;  // [0] bv tests, [1] tri tests
    uint32_t* spill;            // [num_warps][3][spill_cap]
    int32_t spill_cap;
    int32_t* err_flag;
};

// ------------------------------------------------------------------------------------------------ FP64 SE(3)
struct Xf { double q0, q1, q2, q3, tx, ty, tz; };

__device__ __forceinline__ void quat_normalize(double& a, double& b, double& c, double& d)
{
    double f = a * a + b * b + c * c + d * d;
    const double eps = 2.220446049250313e-16;
    if (f < 1.0 - eps || f > 1.0 + eps) {
        f = sqrt(f);
        a /= f; b /= f; c /= f; d /= f;
    }
}

__device__ __forceinline__ void quat_rotate(const Xf& x, double rx, double ry, double rz, double& ox, double& oy, double& oz)
{
    double xx = 2 * x.q1 * x.q1, xy = 2 * x.q1 * x.q2, xz = 2 * x.q1 * x.q3, xw = 2 * x.q1 * x.q0;
    double yy = 2 * x.q2 * x.q2, yz = 2 * x.q2 * x.q3, yw = 2 * x.q2 * x.q0;
    double zz = 2 * x.q3 * x.q3, zw = 2 * x.q3 * x.q0;
    ox = (1 - yy - zz) * rx + (xy - zw) * ry + (xz + yw) * rz;
    oy = (xy + zw) * rx + (1 - xx - zz) * ry + (yz - xw) * rz;
    oz = (xz - yw) * rx + (yz + xw) * ry + (1 - xx - yy) * rz;
}

__device__ __forceinline__ Xf xf_mul(const Xf& a, const Xf& b)     // geometry.h:466-480
{
    Xf r;
    double x, y, z;
    quat_rotate(a, b.tx, b.ty, b.tz, x, y, z);
    r.tx = a.tx + x; r.ty = a.ty + y; r.tz = a.tz + z;
    r.q0 = a.q0 * b.q0 - a.q1 * b.q1 - a.q2 * b.q2 - a.q3 * b.q3;
    r.q1 = a.q0 * b.q1 + a.q1 * b.q0 + a.q2 * b.q3 - a.q3 * b.q2;
    r.q2 = a.q0 * b.q2 + a.q2 * b.q0 + a.q3 * b.q1 - a.q1 * b.q3;
    r.q3 = a.q0 * b.q3 + a.q3 * b.q0 + a.q1 * b.q2 - a.q2 * b.q1;
    quat_normalize(r.q0, r.q1, r.q2, r.q3);
    return r;
}

__device__ __forceinline__ Xf xf_load(const double* p)
{
    Xf r; r.q0 = p[0]; r.q1 = p[1]; r.q2 = p[2]; r.q3 = p[3]; r.tx = p[4]; r.ty = p[5]; r.tz = p[6];
    return r;
}
__device__ __forceinline__ void xf_store(double* p, const Xf& r)
{
    p[0] = r.q0; p[1] = r.q1; p[2] = r.q2; p[3] = r.q3; p[4] = r.tx; p[5] = r.ty; p[6] = r.tz; p[7] = 0.0;
}

// ------------------------------------------------------------------------------------------------ FP32 tests
// 15-axis SAT of box A {ca, ea} (identity frame) against box B {cb, eb} placed by (R, t): x_A = R x_B + t.
// |R| entries are fattened by 1e-6 (OBB_Disjoint.h:67-78); `<=` keeps touching boxes overlapping.
__device__ __forceinline__ bool box_overlap(const float* __restrict__ R, const float* __restrict__ t,
                                            float cax, float cay, float caz, float eax, float eay, float eaz,
                                            float cbx, float cby, float cbz, float ebx, float eby, float ebz)
{
    const float reps = 1e-6f;
    float T0 = fmaf(R[0], cbx, fmaf(R[1], cby, fmaf(R[2], cbz, t[0]))) - cax;
    float T1 = fmaf(R[3], cbx, fmaf(R[4], cby, fmaf(R[5], cbz, t[1]))) - cay;
    float T2 = fmaf(R[6], cbx, fmaf(R[7], cby, fmaf(R[8], cbz, t[2]))) - caz;
    float f00 = fabsf(R[0]) + reps, f01 = fabsf(R[1]) + reps, f02 = fabsf(R[2]) + reps;
    float f10 = fabsf(R[3]) + reps, f11 = fabsf(R[4]) + reps, f12 = fabsf(R[5]) + reps;
    float f20 = fabsf(R[6]) + reps, f21 = fabsf(R[7]) + reps, f22 = fabsf(R[8]) + reps;
    // A's face axes
    if (fabsf(T0) > eax + ebx * f00 + eby * f01 + ebz * f02) return false;
    if (fabsf(T1) > eay + ebx * f10 + eby * f11 + ebz * f12) return false;
    if (fabsf(T2) > eaz + ebx * f20 + eby * f21 + ebz * f22) return false;
    // B's face axes
    if (fabsf(T0 * R[0] + T1 * R[3] + T2 * R[6]) > ebx + eax * f00 + eay * f10 + eaz * f20) return false;
    if (fabsf(T0 * R[1] + T1 * R[4] + T2 * R[7]) > eby + eax * f01 + eay * f11 + eaz * f21) return false;
    if (fabsf(T0 * R[2] + T1 * R[5] + T2 * R[8]) > ebz + eax * f02 + eay * f12 + eaz * f22) return false;
    // edge x edge
    if (fabsf(T2 * R[3] - T1 * R[6]) > eay * f20 + eaz * f10 + eby * f02 + ebz * f01) return false;
    if (fabsf(T2 * R[4] - T1 * R[7]) > eay * f21 + eaz * f11 + ebx * f02 + ebz * f00) return false;
    if (fabsf(T2 * R[5] - T1 * R[8]) > eay * f22 + eaz * f12 + ebx * f01 + eby * f00) return false;
    if (fabsf(T0 * R[6] - T2 * R[0]) > eax * f20 + eaz * f00 + eby * f12 + ebz * f11) return false;
    if (fabsf(T0 * R[7] - T2 * R[1]) > eax * f21 + eaz * f01 + ebx * f12 + ebz * f10) return false;
    if (fabsf(T0 * R[8] - T2 * R[2]) > eax * f22 + eaz * f02 + ebx * f11 + eby * f10) return false;
    if (fabsf(T1 * R[0] - T0 * R[3]) > eax * f10 + eay * f00 + eby * f22 + ebz * f21) return false;
    if (fabsf(T1 * R[1] - T0 * R[4]) > eax * f11 + eay * f01 + ebx * f22 + ebz * f20) return false;
    if (fabsf(T1 * R[2] - T0 * R[5]) > eax * f12 + eay * f02 + ebx * f21 + eby * f20) return false;
    return true;
}

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// project6 (pqp/PQP.cpp:350-369): true when `ax` does NOT separate.  p0 is the origin after the shift.
// With eps > 0 separation needs a gap larger than eps * |ax| (conservative); eps = 0 is the reference's
// strict test.
__device__ __forceinline__ bool axis_overlaps(V3 ax, V3 p1, V3 p2, V3 q0, V3 q1, V3 q2, float eps2)
{
    float P1 = dot(ax, p1), P2 = dot(ax, p2);
    float Q0 = dot(ax, q0), Q1 = dot(ax, q1), Q2 = dot(ax, q2);
    float mx1 = fmaxf(0.f, fmaxf(P1, P2)), mn1 = fminf(0.f, fminf(P1, P2));
    float mx2 = fmaxf(Q0, fmaxf(Q1, Q2)), mn2 = fminf(Q0, fminf(Q1, Q2));
    float gap = fmaxf(mn1 - mx2, mn2 - mx1);
    if (gap <= 0.f) return true;
    if (eps2 == 0.f) return false;
    return gap * gap <= eps2 * dot(ax, ax);
}

// TriContact (pqp/PQP.cpp:375-459): 17 axes in the reference's order n1, m1, e_i x f_j, e_i x n1, f_i x m1.
__device__ __forceinline__ bool tri_contact(V3 P0, V3 P1, V3 P2, V3 Q0, V3 Q1, V3 Q2, float eps2)
{
    V3 p1 = P1 - P0, p2 = P2 - P0;
    V3 q0 = Q0 - P0, q1 = Q1 - P0, q2 = Q2 - P0;
    V3 e1 = p1, e2 = p2 - p1, e3 = v3(-p2.x, -p2.y, -p2.z);
    V3 f1 = q1 - q0, f2 = q2 - q1, f3 = q0 - q2;
    V3 n1 = cross(e1, e2);
    if (!axis_overlaps(n1, p1, p2, q0, q1, q2, eps2)) return false;
    V3 m1 = cross(f1, f2);
    if (!axis_overlaps(m1, p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e1, f1), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e1, f2), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e1, f3), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e2, f1), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e2, f2), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e2, f3), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e3, f1), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e3, f2), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e3, f3), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e1, n1), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e2, n1), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(e3, n1), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(f1, m1), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(f2, m1), p1, p2, q0, q1, q2, eps2)) return false;
    if (!axis_overlaps(cross(f3, m1), p1, p2, q0, q1, q2, eps2)) return false;
    return true;
}

// ------------------------------------------------------------------------------------------------ shared memory
// Per-warp layout (bytes):  xf32 [nq*48] | slot_rt [kNumSlots*48] | slot_pair [kNumSlots*4] | slot_hit [16]
//                           | scratch: max(stack 3*kStackCap*4, FK (nq + nsteps)*64)
__host__ __device__ inline size_t warp_smem_bytes(int nq, int nsteps)
{
    size_t fk = (size_t)(nq + nsteps) * 64;
    size_t st = (size_t)3 * kStackCap * 4;
    size_t scratch = fk > st ? fk : st;
    size_t b = (size_t)nq * 48 + (size_t)kNumSlots * 48 + (size_t)kNumSlots * 4 + 16 + scratch;
    return (b + 15) & ~(size_t)15;
}

struct WarpState {
    float* xf32;          // [nq][12]
    float* slot_rt;       // [kNumSlots][12]
    int32_t* slot_pair;   // [kNumSlots] original pair index
    uint32_t* slot_hit;   // [2]
    uint32_t* st_a;       // [kStackCap]
    uint32_t* st_b;
    uint32_t* st_s;
    double* fk_xf;        // [nq][8]     (aliases the stack)
    double* fk_m;         // [nsteps][8]
    uint32_t* spill;      // global [3][spill_cap]
    int sp;
    int spill_count;
    int best_env, best_self;
    unsigned long long n_bv, n_tri;
};

__device__ __forceinline__ void relative_pose(const float* A, const float* B, float* R, float* t)
{
    // x_A = Ra^T (Rb x_B + tb - ta)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            R[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
    float dx = B[9] - A[9], dy = B[10] - A[10], dz = B[11] - A[11];
    t[0] = A[0] * dx + A[3] * dy + A[6] * dz;
    t[1] = A[1] * dx + A[4] * dy + A[7] * dz;
    t[2] = A[2] * dx + A[5] * dy + A[8] * dz;
}

// Move the oldest kSpillChunk items to global memory and slide the rest down.
__device__ __noinline__ void stack_spill(const Params& P, WarpState& w, int lane)
{
    if (w.spill_count + kSpillChunk > P.spill_cap) {
        if (lane == 0) atomicExch(P.err_flag, 1);
        w.sp = 0;               // give up on this configuration; the host reports the error
        __syncwarp();
        return;
    }
    uint32_t* ga = w.spill + w.spill_count;
    uint32_t* gb = ga + P.spill_cap;
    uint32_t* gs = gb + P.spill_cap;
    for (int i = lane; i < kSpillChunk; i += 32) { ga[i] = w.st_a[i]; gb[i] = w.st_b[i]; gs[i] = w.st_s[i]; }
    w.spill_count += kSpillChunk;
    int rest = w.sp - kSpillChunk;
    __syncwarp();
    for (int base = 0; base < rest; base += 32) {
        int i = base + lane;
        uint32_t a = 0, b = 0, s = 0;
        if (i < rest) { a = w.st_a[kSpillChunk + i]; b = w.st_b[kSpillChunk + i]; s = w.st_s[kSpillChunk + i]; }
        __syncwarp();
        if (i < rest) { w.st_a[i] = a; w.st_b[i] = b; w.st_s[i] = s; }
        __syncwarp();
    }
    w.sp = rest;
}

__device__ __noinline__ void stack_refill(const Params& P, WarpState& w, int lane)
{
    int m = w.spill_count < kSpillChunk ? w.spill_count : kSpillChunk;
    const uint32_t* ga = w.spill + (w.spill_count - m);
    const uint32_t* gb = ga + P.spill_cap;
    const uint32_t* gs = gb + P.spill_cap;
    for (int i = lane; i < m; i += 32) { w.st_a[i] = ga[i]; w.st_b[i] = gb[i]; w.st_s[i] = gs[i]; }
    w.spill_count -= m;
    w.sp = m;
    __syncwarp();
}

// Drain the item stack: the warp-cooperative dual-tree descent + triangle tests.
__device__ __forceinline__ void traverse(const Params& P, WarpState& w, int lane)
{
    const bool all_pairs = (P.flags & kFlagAllPairs) != 0;
    const float eps2 = P.tri_eps * P.tri_eps;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (;;) {
        if (w.sp == 0) {
            if (w.spill_count > 0) stack_refill(P, w, lane); else break;
        }
        const int n = w.sp < 32 ? w.sp : 32;
        bool valid = lane < n;
        uint32_t a = 0, b = 0, s = 0;
        if (valid) { int idx = w.sp - 1 - lane; a = w.st_a[idx]; b = w.st_b[idx]; s = w.st_s[idx]; }
        w.sp -= n;
        __syncwarp();
        int pidx = 0;
        if (valid) {
            pidx = w.slot_pair[s];
            bool dead = (w.slot_hit[s >> 5] >> (s & 31)) & 1u;
            if (!all_pairs) dead |= (pidx < P.nenv_total) ? (pidx > w.best_env) : (pidx > w.best_self);
            valid = !dead;
        }
        bool push = false, hit = false;
        uint32_t c0a = 0, c0b = 0, c1a = 0, c1b = 0;
        if (valid) {
            const float4 a0 = __ldg(&P.nodes[2 * a]), a1 = __ldg(&P.nodes[2 * a + 1]);
            const float4 b0 = __ldg(&P.nodes[2 * b]), b1 = __ldg(&P.nodes[2 * b + 1]);
            const float* rt = w.slot_rt + 12 * s;
            float R[9], t[3];
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = rt[k];
            t[0] = rt[9]; t[1] = rt[10]; t[2] = rt[11];
            if (P.flags & kFlagCounters) w.n_bv++;
            if (box_overlap(R, t, a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, b0.x, b0.y, b0.z, b1.x, b1.y, b1.z)) {
                const int la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
                if (la < 0 && lb < 0) {
                    if (P.flags & kFlagCounters) w.n_tri++;
                    const float4* ta = P.tris + 3 * (size_t)(~la);
                    const float4* tb = P.tris + 3 * (size_t)(~lb);
                    const float4 p0 = __ldg(ta), p1 = __ldg(ta + 1), p2 = __ldg(ta + 2);
                    const float4 u0 = __ldg(tb), u1 = __ldg(tb + 1), u2 = __ldg(tb + 2);
                    V3 Q0 = v3(fmaf(R[0], u0.x, fmaf(R[1], u0.y, fmaf(R[2], u0.z, t[0]))),
                               fmaf(R[3], u0.x, fmaf(R[4], u0.y, fmaf(R[5], u0.z, t[1]))),
                               fmaf(R[6], u0.x, fmaf(R[7], u0.y, fmaf(R[8], u0.z, t[2]))));
                    V3 Q1 = v3(fmaf(R[0], u1.x, fmaf(R[1], u1.y, fmaf(R[2], u1.z, t[0]))),
                               fmaf(R[3], u1.x, fmaf(R[4], u1.y, fmaf(R[5], u1.z, t[1]))),
                               fmaf(R[6], u1.x, fmaf(R[7], u1.y, fmaf(R[8], u1.z, t[2]))));
                    V3 Q2 = v3(fmaf(R[0], u2.x, fmaf(R[1], u2.y, fmaf(R[2], u2.z, t[0]))),
                               fmaf(R[3], u2.x, fmaf(R[4], u2.y, fmaf(R[5], u2.z, t[1]))),
                               fmaf(R[6], u2.x, fmaf(R[7], u2.y, fmaf(R[8], u2.z, t[2]))));
                    hit = tri_contact(v3(p0.x, p0.y, p0.z), v3(p1.x, p1.y, p1.z), v3(p2.x, p2.y, p2.z), Q0, Q1, Q2, eps2);
                } else {
                    // descend the larger box (PQP.cpp:539-544 compares BV sizes; any rule is verdict-neutral)
                    const float sa = a1.x * a1.x + a1.y * a1.y + a1.z * a1.z;
                    const float sb = b1.x * b1.x + b1.y * b1.y + b1.z * b1.z;
                    push = true;
                    if (lb < 0 || (la >= 0 && sa > sb)) {
                        c0a = (uint32_t)la; c1a = (uint32_t)__float_as_int(a1.w); c0b = b; c1b = b;
                    } else {
                        c0a = a; c1a = a; c0b = (uint32_t)lb; c1b = (uint32_t)__float_as_int(b1.w);
                    }
                }
            }
        }
        const unsigned hm = __ballot_sync(kFull, hit);
        if (hm) {
            if (hit) atomicOr(&w.slot_hit[s >> 5], 1u << (s & 31));
            int pe = (hit && pidx < P.nenv_total) ? pidx : 0x7fffffff;
            int ps = (hit && pidx >= P.nenv_total) ? pidx : 0x7fffffff;
            pe = __reduce_min_sync(kFull, pe);
            ps = __reduce_min_sync(kFull, ps);
            if (pe < w.best_env) w.best_env = pe;
            if (ps < w.best_self) w.best_self = ps;
        }
        const unsigned pm = __ballot_sync(kFull, push);
        if (push) {
            const int off = w.sp + 2 * __popc(pm & lt_mask);
            w.st_a[off] = c1a; w.st_b[off] = c1b; w.st_s[off] = s;
            w.st_a[off + 1] = c0a; w.st_b[off + 1] = c0b; w.st_s[off + 1] = s;
        }
        w.sp += 2 * __popc(pm);
        __syncwarp();
        if (w.sp > kStackCap - 64) stack_spill(P, w, lane);
    }
}

// Write the hit bits of the current slots into the global pair bitmask (all-pairs mode).
__device__ __forceinline__ void flush_slots(const Params& P, WarpState& w, int lane, int nslots_used, uint32_t* mask_row)
{
    if (mask_row) {
        for (int s = lane; s < nslots_used; s += 32) {
            if ((w.slot_hit[s >> 5] >> (s & 31)) & 1u) {
                int p = w.slot_pair[s];
                atomicOr(&mask_row[p >> 5], 1u << (p & 31));
            }
        }
    }
    __syncwarp();
    if (lane < 2) w.slot_hit[lane] = 0;
    __syncwarp();
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
cc_check_kernel(const __grid_constant__ Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const size_t wbytes = warp_smem_bytes(P.nq, P.nsteps);
    unsigned char* base = smem_raw + wib * wbytes;
    WarpState w;
    w.xf32 = (float*)base;
    w.slot_rt = w.xf32 + (size_t)P.nq * 12;
    w.slot_pair = (int32_t*)(w.slot_rt + kNumSlots * 12);
    w.slot_hit = (uint32_t*)(w.slot_pair + kNumSlots);
    unsigned char* scratch = (unsigned char*)(w.slot_hit + 4);
    w.st_a = (uint32_t*)scratch;
    w.st_b = w.st_a + kStackCap;
    w.st_s = w.st_b + kStackCap;
    w.fk_xf = (double*)scratch;
    w.fk_m = w.fk_xf + (size_t)P.nq * 8;
    const int gwarp = blockIdx.x * kWarpsPerBlock + wib;
    w.spill = P.spill + (size_t)gwarp * 3 * P.spill_cap;
    w.n_bv = 0; w.n_tri = 0;
    const bool all_pairs = (P.flags & kFlagAllPairs) != 0;

    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(P.work_counter, 1);
        c = __shfl_sync(kFull, c, 0);
        if (c >= P.n) break;

        // ---------------------------------------------------------------- 1. link poses (FP64)
        if (P.flags & kFlagPoseMode) {
            const double* src = P.q + (size_t)c * P.nq * 8;
            for (int i = lane; i < P.nq * 8; i += 32) w.fk_xf[i] = src[i];
        } else {
            const double* qrow = P.q + (size_t)c * P.ncols;
            for (int s = lane; s < P.nsteps; s += 32) {
                const FkStep& st = P.steps[s];
                Xf L = xf_load(st.L);
                Xf M;
                if (st.flags & 1) {
                    M = L;                                   // static joint: right already folded (kinbodyjoint.cpp:1527-1551)
                } else {
                    double v = st.src_col >= 0 ? qrow[st.src_col] : st.src_const;
                    if (!(st.flags & 8)) {                   // SetDOFValues clamps non-circular DOFs (kinbody.cpp:2149-2224)
                        if (v < st.src_lower) v = st.src_lower; else if (v > st.src_upper) v = st.src_upper;
                    }
                    if (st.flags & 2) {                      // mimic (kinbody.cpp:2308-2386)
                        double ev = st.a * v + st.b;
                        v = ev;
                        if (!(st.flags & 4)) {
                            if (ev < st.lower) { if (ev >= st.lower - 1e-7) v = st.lower; else if (ev < st.lower - 5e-4) v = st.lower; }
                            else if (ev > st.upper) { if (ev <= st.upper + 1e-7) v = st.upper; else if (ev > st.upper + 5e-4) v = st.upper; }
                        }
                    }
                    Xf J; J.q0 = 1; J.q1 = J.q2 = J.q3 = 0; J.tx = J.ty = J.tz = 0;
                    if (st.type == 1) {
                        // quatFromAxisAngle (geometry.h:673-682)
                        double len = sqrt(st.axis[0] * st.axis[0] + st.axis[1] * st.axis[1] + st.axis[2] * st.axis[2]);
                        if (len != 0) {
                            double sn, cs; sincos(v * 0.5, &sn, &cs);
                            sn /= len;
                            J.q0 = cs; J.q1 = st.axis[0] * sn; J.q2 = st.axis[1] * sn; J.q3 = st.axis[2] * sn;
                        }
                    } else {
                        J.tx = st.axis[0] * v; J.ty = st.axis[1] * v; J.tz = st.axis[2] * v;
                    }
                    M = xf_mul(xf_mul(L, J), xf_load(st.R));  // (L * J) * R, kinbody.cpp:2478
                }
                xf_store(w.fk_m + 8 * s, M);
            }
            if (lane < 8) w.fk_xf[lane] = P.base[lane];       // link 0 keeps its pose (kinbody.cpp:2281)
            __syncwarp();
            for (int d = 0; d <= P.max_depth; ++d) {
                const int s0 = P.level_start[d], s1 = P.level_start[d + 1];
                for (int s = s0 + lane; s < s1; s += 32) {
                    const FkStep& st = P.steps[s];
                    Xf T = xf_mul(xf_load(w.fk_xf + 8 * st.parent), xf_load(w.fk_m + 8 * s));
                    xf_store(w.fk_xf + 8 * st.child, T);
                }
                __syncwarp();
            }
        }
        __syncwarp();
        if (P.fk_out) {
            double* dst = P.fk_out + (size_t)c * P.nq * 8;
            for (int i = lane; i < P.nq * 8; i += 32) dst[i] = w.fk_xf[i];
        }
        if (P.flags & kFlagFkOnly) { __syncwarp(); continue; }
        // FP32 R|t, base-centred (matrixFromQuat, geometry.h:779-800)
        for (int l = lane; l < P.nq; l += 32) {
            const double* x = w.fk_xf + 8 * l;
            double q0 = x[0], q1 = x[1], q2 = x[2], q3 = x[3];
            double il = 2.0 / (q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
            double qq1 = il * q1 * q1, qq2 = il * q2 * q2, qq3 = il * q3 * q3;
            float* o = w.xf32 + 12 * l;
            o[0] = (float)(1 - qq2 - qq3);            o[1] = (float)(il * (q1 * q2 - q0 * q3)); o[2] = (float)(il * (q1 * q3 + q0 * q2));
            o[3] = (float)(il * (q1 * q2 + q0 * q3)); o[4] = (float)(1 - qq1 - qq3);            o[5] = (float)(il * (q2 * q3 - q0 * q1));
            o[6] = (float)(il * (q1 * q3 - q0 * q2)); o[7] = (float)(il * (q2 * q3 + q0 * q1)); o[8] = (float)(1 - qq1 - qq2);
            o[9] = (float)(x[4] - P.center[0]); o[10] = (float)(x[5] - P.center[1]); o[11] = (float)(x[6] - P.center[2]);
        }
        __syncwarp();   // from here on the scratch area is the item stack

        // ---------------------------------------------------------------- 2.+3.+4. pairs
        w.sp = 0; w.spill_count = 0;
        w.best_env = P.hoisted_first_env;
        w.best_self = P.hoisted_first_self;
        if (lane < 2) w.slot_hit[lane] = 0;
        uint32_t* mask_row = NULL;
        if (P.pair_mask) {
            mask_row = P.pair_mask + (size_t)c * P.mask_words;
            for (int i = lane; i < P.mask_words; i += 32) mask_row[i] = P.hoisted_mask ? P.hoisted_mask[i] : 0u;
        }
        __syncwarp();
        int nslots_used = 0;
        const unsigned lt_mask = (1u << lane) - 1u;
        for (int basek = 0; basek < P.npairs_dev; basek += 32) {
            const int k = basek + lane;
            bool cand = false;
            float R[9], t[3];
            int orig = 0;
            uint32_t ra = 0, rb = 0;
            if (k < P.npairs_dev) {
                orig = P.pair_orig[k];
                const bool is_env = k < P.nenv_dev;
                bool skip = !(P.flags & (is_env ? kFlagEnv : kFlagSelf));
                if (!all_pairs) skip |= is_env ? (orig > w.best_env) : (orig > w.best_self);
                if (!skip) {
                    const uint32_t pr = P.pairs[k];
                    const int la = pr & 0xffff, lb = pr >> 16;
                    const float* A = w.xf32 + 12 * la;
                    float Bx[12];
                    if (lb < P.nq) {
#pragma unroll
                        for (int i = 0; i < 12; ++i) Bx[i] = w.xf32[12 * lb + i];
                    } else {
                        const float4* e = (const float4*)(P.env_xf + 12 * (size_t)(lb - P.nq));
                        float4 e0 = __ldg(e), e1 = __ldg(e + 1), e2 = __ldg(e + 2);
                        Bx[0] = e0.x; Bx[1] = e0.y; Bx[2] = e0.z; Bx[3] = e0.w; Bx[4] = e1.x; Bx[5] = e1.y; Bx[6] = e1.z; Bx[7] = e1.w;
                        Bx[8] = e2.x; Bx[9] = e2.y; Bx[10] = e2.z; Bx[11] = e2.w;
                    }
                    relative_pose(A, Bx, R, t);
                    ra = (uint32_t)P.slot_root[la]; rb = (uint32_t)P.slot_root[lb];
                    const float4 a0 = __ldg(&P.nodes[2 * ra]), a1 = __ldg(&P.nodes[2 * ra + 1]);
                    const float4 b0 = __ldg(&P.nodes[2 * rb]), b1 = __ldg(&P.nodes[2 * rb + 1]);
                    if (P.flags & kFlagCounters) w.n_bv++;
                    cand = box_overlap(R, t, a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, b0.x, b0.y, b0.z, b1.x, b1.y, b1.z);
                }
            }
            const unsigned cm = __ballot_sync(kFull, cand);
            if (cm) {
                const int cnt = __popc(cm);
                if (cand) {
                    const int rank = __popc(cm & lt_mask);
                    const int s = nslots_used + rank;
                    float* rt = w.slot_rt + 12 * s;
#pragma unroll
                    for (int i = 0; i < 9; ++i) rt[i] = R[i];
                    rt[9] = t[0]; rt[10] = t[1]; rt[11] = t[2];
                    w.slot_pair[s] = orig;
                    // reversed inside the chunk so the lowest pair index is popped first
                    const int pos = w.sp + (cnt - 1 - rank);
                    w.st_a[pos] = ra; w.st_b[pos] = rb; w.st_s[pos] = (uint32_t)s;
                }
                nslots_used += cnt;
                w.sp += cnt;
            }
            __syncwarp();
            const bool last = basek + 32 >= P.npairs_dev;
            if (nslots_used > kNumSlots - 32 || (last && nslots_used > 0)) {
                traverse(P, w, lane);
                flush_slots(P, w, lane, nslots_used, all_pairs ? mask_row : NULL);
                nslots_used = 0;
            }
        }

        // ---------------------------------------------------------------- 5. verdict
        if (lane == 0) {
            int be = w.best_env, bs = w.best_self;
            if (all_pairs && mask_row) {
                // complete masks: derive the first pairs from the bits
                be = 0x7fffffff; bs = 0x7fffffff;
                const int total = P.mask_words * 32;
                for (int wd = 0; wd < P.mask_words; ++wd) {
                    uint32_t m = mask_row[wd];
                    while (m) {
                        int bit = __ffs(m) - 1; m &= m - 1;
                        int p = wd * 32 + bit;
                        if (p < P.nenv_total) { if (p < be) be = p; } else { if (p < bs) bs = p; }
                    }
                    (void)total;
                }
            }
            uint32_t v = 0;
            if ((P.flags & kFlagEnv) && be != 0x7fffffff) v |= 1u;
            if ((P.flags & kFlagSelf) && bs != 0x7fffffff) v |= 2u;
            if (P.verdict) P.verdict[c] = v;
            if (P.first_env) P.first_env[c] = (v & 1u) ? be : -1;
            if (P.first_self) P.first_self[c] = (v & 2u) ? bs : -1;
        }
        __syncwarp();
    }
    if (P.flags & kFlagCounters) {
        unsigned long long bv = w.n_bv, tr = w.n_tri;
        for (int o = 16; o > 0; o >>= 1) { bv += __shfl_xor_sync(kFull, bv, o); tr += __shfl_xor_sync(kFull, tr, o); }
        if (lane == 0) { atomicAdd(&P.stats[0], bv); atomicAdd(&P.stats[1], tr); }
    }
}

}  // namespace ccdev
