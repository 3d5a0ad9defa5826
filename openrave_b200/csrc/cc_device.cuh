// cc_device.cuh — device side of the batched collision checker (sm_100a only).
//
// Three kernels per chunk of configurations, all hand-written for B200:
//
//   cc_fk_kernel        ONE THREAD per configuration.  The FK chain program is uniform across the warp, so
//                       KinBody::SetDOFValues' joint loop (kinbody.cpp:2286-2481) runs divergence-free in FP64 with
//                       the running transform in registers (the chain is walked depth-first and pruned to the links
//                       the bound pairs need; only branch-point links are parked in a coalesced SoA scratch).  Every
//                       needed link pose is emitted as FP32 R|t (base-centred) plus its world AABB.
//   cc_fk2_kernel       the same arithmetic in two phases (varying steps per configuration, then one thread per
//                       (configuration, link) with coalesced stores): chosen per plan when rigidly attached links
//                       dominate, and for pose mode.
//   cc_broad_kernel     world-AABB test of every (configuration, pair) + certificate-cell lookup -> one candidate bit per
//                       pair: a block per 32 configurations, their boxes staged in shared memory as 15-bit fixed point,
//                       a warp per word of 32 pairs.
//   cc_traverse_kernel  persistent, ONE WARP per configuration: candidate bits (or, for small split batches, an in-warp
//                       link-AABB x {static-AABB, link-AABB} broadphase over the bound pair table in reference order)
//                       -> dense batches of 32 candidate root pairs (relative pose + SAT of the root boxes) ->
//                       warp-cooperative dual-BVH descent.  A stack item names the two CHILDREN of the node being split
//                       plus the other node, so one pop is three 256-bit node loads and two independent box tests per
//                       lane; leaf x leaf hits are parked in a triangle queue and tested 32 at a time with PQP's 17-axis
//                       triangle SAT (pqp/PQP.cpp:350-459) in FP32.  All warp state lives in registers; the stack, the
//                       triangle queues and the per-pair relative poses live in 7.5 KB of shared memory per warp.
//
// Conservative pruning: node boxes are inflated at build time and |R| entries get +1e-6 like
// OBB_Disjoint.h:67-78, so FP32 rounding can never prune a true contact; the verdict is the OR of the
// triangle tests of unpruned pairs, exactly like the reference's (BVH-shape independent).
#pragma once
#include <cuda_runtime.h>
#ifndef CC_EDGE_AXES
#define CC_EDGE_AXES 0
#endif
#ifndef CC_TRI_T1
#define CC_TRI_T1 16     // queued leaf pairs that trigger a face-normal pass (contacts found early prune the most work:
                         // measured 16/8 > 32/16 > 32/32 on all workloads, profiles/r01c_ab_tri.log)
#endif
#ifndef CC_TRI_T2
#define CC_TRI_T2 8      // survivors that trigger an edge-axes pass
#endif
#include <stdint.h>
#include "cc_epsilons.h"

namespace ccdev {

#ifndef CC_WARPS_PER_BLOCK
#define CC_WARPS_PER_BLOCK 2     // 2 warps per block, 14 blocks per SM: a block's slot frees as soon as its two warps run dry (+0.1 .. 0.4 % over 4 x 7, profiles/r02B_ab.log)
#endif
constexpr int kWarpsPerBlock = CC_WARPS_PER_BLOCK;
#ifndef CC_MIN_BLOCKS
#define CC_MIN_BLOCKS 14
#endif
constexpr int kMinBlocksPerSM = CC_MIN_BLOCKS;   // shared memory allows 28 warps per SM (14 blocks of 2): the register cap follows that (73), not 32 warps (64)
constexpr int kFkThreads = 128;
#ifndef CC_STACK_CAP
#define CC_STACK_CAP 320
#endif
#ifndef CC_PUSH_REVERSED
#define CC_PUSH_REVERSED 1     // 1: the outputs of the lane that popped the top item go back on top (strict pair-order priority: +18 .. 37 %, profiles/r01I_ab_push_order.log)
#endif
#ifndef CC_LD256
#define CC_LD256 1           // 256-bit loads (LDG.E.256, sm_100) for the 32-byte BVH nodes and boxes (0: pairs of 128-bit loads, A/B)
#endif
#ifndef CC_TRIQ_CAP
#define CC_TRIQ_CAP 96
#endif
constexpr int kStackCap = CC_STACK_CAP;   // items of shared-memory LIFO per warp
constexpr int kSpillChunk = 128;     // items moved to / from the global spill area at a time
constexpr int kNumSlots = 32;        // root pairs traversed together
constexpr int kTriQCap = CC_TRIQ_CAP;         // pending leaf x leaf tests (drained 32 at a time)
constexpr int kTriQ2Cap = 64;        // leaf x leaf tests that survived the two face-normal axes (15 axes left, drained 32 at a time)
constexpr int kBroadPerLane = 4;     // broadphase pairs per lane per step (independent loads in flight)
constexpr int kCandCap = 32 * kBroadPerLane + 32;   // pending candidate pair indices (a step only runs with < 32 pending)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kNone = 0x7fffffff;

// flags (low bits mirror cc_batch_flags)
constexpr int kFlagEnv = 1, kFlagSelf = 2, kFlagAllPairs = 4, kFlagPoseMode = 8, kFlagFkOnly = 16, kFlagCounters = 32;
constexpr int kFlagFirstHit = 128;     // verdict only: an env hit ends the configuration, a self hit ends the self pairs
constexpr int kFlagInitOutputs = 64;   // split traversal: the FK kernel initialises the outputs the traversal combines into

// The chain program (compiled per plan by compile_fk_program, cudacollision.cu).  Only joints whose value depends on an ACTIVE DOF
// are walked ("varying" steps: T_child = T_parent * L * J(q) * R, kinbody.cpp:2478); everything rigidly attached to such a joint's
// child link — static joints, joints of inactive DOFs, grabbed links — is folded on the host into one constant per link and
// becomes a "constant" step right after it: T_link = T_anchor * L, which leaves the running transform alone.  The left
// transform of a varying step likewise absorbs the constant path from the previous varying joint's child to its own parent.
struct FkStep {                 // 256 bytes
    int32_t type;               // 1 revolute, 2 prismatic
    int32_t parent, child;      // links (parent: the link whose pose the step starts from)
    int32_t src_col;            // column of q feeding this joint (own DOF or the mimic source); -1: constant
    int32_t flags;              // bit0 static / constant joint (only L is used), bit1 mimic, bit2 own-circular, bit3 source-circular,
                                // bit4 constant step: the child's pose is emitted as running * L, the running transform is kept
    int32_t load_parent;        // -2: parent pose is the running transform, -1: link 0's pose, >=0: parked pose index
    int32_t save;               // >=0: park this child's pose (a later step branches off it)
    int32_t pose_in;            // two-phase FK kernel: index (0 = link 0, 1 + j = child of the j-th varying step) of the pose this step starts from
    double src_const;           // value when src_col < 0
    double src_lower, src_upper;// limits of the source DOF (KinBody::SetDOFValues clamps DOF values first)
    double a, b;                // mimic: value = a*src + b
    double lower, upper;        // own limits (mimic clamp rule)
    double L[8];                // left transform  {qw qx qy qz tx ty tz 0}
    double R[8];                // right transform
    double axis[4];
    int32_t pose_out;           // two-phase FK kernel: where a varying step stores its child's pose (-1: constant step)
    int32_t pad2;
};
static_assert(sizeof(FkStep) == 256, "FkStep layout");

struct PairTable;

struct Params {
    // scene (device pointers)
    const float4* nodes;        // 2 x float4 per node: {cx cy cz | left} {ex ey ez | right}
    const float4* tris;         // 3 x float4 per triangle (link frame)
    // plan = (pair list, slot space).  Slots [0, ndyn) are query links whose pose changes per configuration (scratch),
    // slots [ndyn, ndyn + nstatic) are everything static for this plan: environment links and query links no active DOF moves.
    const int32_t* slot_root;   // [ndyn + nstatic] root node of every slot
    const float* st_xf;         // [nstatic][12] FP32 R(9, row major)|t(3), base-centred
    const float* st_aabb;       // [nstatic][8]  world AABB {cx cy cz 0 ex ey ez 0}, base-centred
    const float* link_box;      // [ndyn][8] link-frame root box {cx cy cz ex ey ez 1 0}
    const int32_t* emit_index;  // [nq] link -> dynamic slot, -1: pose not needed
    const uint32_t* pairs;      // device pair list: slot_a | (slot_b << 16)
    const int32_t* pair_orig;   // original (reference-order) index of each device pair, increasing within env / self
    const FkStep* steps;        // depth-first order, pruned to the ancestors of the dynamic slots
    const uint32_t* hoisted_mask;
    int32_t nq, ndyn, nstatic, npairs_dev, nenv_dev, nenv_total, nsteps, ncols;
    int32_t npose;              // poses the two-phase FK kernel keeps per configuration: link 0 + one per varying step
    int32_t hoisted_first_env, hoisted_first_self, mask_words, flags;
    double base[8];             // pose of link 0 (kept, kinbody.cpp:2281)
    // affine active DOFs (RobotBase::SetActiveDOFValues, robot.cpp:1193-1225): the columns [affine_col, ..) of q hold X | Y | Z and one
    // of RotationAxis (angle) | Rotation3D (axis * angle) | RotationQuat (4, normalised, left-multiplied by affine_qstart) and
    // overwrite those components of link 0's pose per configuration (RaveGetTransformFromAffineDOFValues, libopenrave.cpp:1868-1924)
    int32_t affine_mask, affine_col;
    double affine_axis[4];
    double affine_qstart[4];
    double center[3];           // subtracted from every translation before the FP32 conversion
    float tri_eps;              // conservative margin of the triangle test (metres), 0 = reference's strict test
    // io (already offset to this chunk)
    const double* q;            // [n][ncols]   (or [n][nq][8] in pose mode)
    int32_t n;
    uint32_t* verdict;
    int32_t* first_env;
    int32_t* first_self;
    uint32_t* pair_mask;
    double* fk_out;             // [n][nq][8] or NULL
    uint32_t* work_out;         // [n] or NULL: traversal iterations spent on each configuration (diagnostics)
    // per-chunk scratch
    float* xf32;                // [cap][ndyn][12]
    float* aabb;                // [cap][ndyn][8]
    double* fksave;             // [nsave][7][cap]
    int32_t* counters;          // [0] traversal work counter
    int32_t* err_flag;          // set when a traversal overflows its spill area
    int32_t cap;
    unsigned long long* stats;  // [0] bv tests, [1] tri tests
    uint2* spill;               // [num_warps][spill_cap]
    int32_t spill_cap;
    // split traversal (small batches): every configuration becomes `split` work items, item `part` owning the pairs
    // [part_begin[part], part_begin[part + 1]) of `pairs` / `pair_orig` (round-robin blocks of the reference order, env pairs
    // first: part_nenv[part] of them); results are combined with atomics into outputs the FK kernel initialised
    int32_t split;
    const int32_t* part_begin;  // [split + 1]
    const int32_t* part_nenv;   // [split]
    // certified free-space tables (FK mode only): pair_table[k] >= 0 names a table over the 1 or 2 active DOFs the pair's
    // relative pose depends on; a set bit = every configuration of that cell is at least the contact band away from contact
    // (distance at the cell centre > the largest displacement any point can make inside the cell), so the pair is skipped
    // all-contacts flavour (CO_Contacts on the report path): every intersecting triangle pair is appended as
    // {configuration, reference pair index, triangle of slot a, triangle of slot b} (global triangle indices)
    uint4* contact_out;
    uint32_t* contact_count;    // [0] number of records produced (may exceed contact_cap: the rest is dropped)
    int32_t contact_cap;
    const int32_t* pair_table;  // [npairs_dev] (same order as `pairs`), or NULL
    const struct PairTable* tables;
    const uint32_t* table_bits;
    // broadphase candidates (full batches): bit k of row c = device pair k passed the world-AABB test (and no certificate covers
    // the configuration's cell); written by cc_broad_kernel, one thread per configuration, read by the traversal's kCand flavour
    uint32_t* cand;             // [cap][cand_words]
    int32_t cand_words;
};

struct PairTable {
    int32_t col0, col1;         // columns of q (col1 = -1: one DOF)
    int32_t n0, n1;             // cells per DOF (n1 = 1 for one DOF)
    int32_t circ0, circ1;       // circular DOFs are not clamped by SetDOFValues: 1 = outside the range the table does not apply,
                                // 2 = the range is the full circle, the value is wrapped into it
    uint32_t bit_base;          // first bit of this table in table_bits
    int32_t pad;
    double lo0, hi0, inv0;      // cell = (int)((clamp(v) - lo) * inv), the last cell closed
    double lo1, hi1, inv1;
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
    bool sep = fabsf(T0) > eax + ebx * f00 + eby * f01 + ebz * f02;
    sep |= fabsf(T1) > eay + ebx * f10 + eby * f11 + ebz * f12;
    sep |= fabsf(T2) > eaz + ebx * f20 + eby * f21 + ebz * f22;
    // B's face axes
    sep |= fabsf(T0 * R[0] + T1 * R[3] + T2 * R[6]) > ebx + eax * f00 + eay * f10 + eaz * f20;
    sep |= fabsf(T0 * R[1] + T1 * R[4] + T2 * R[7]) > eby + eax * f01 + eay * f11 + eaz * f21;
    sep |= fabsf(T0 * R[2] + T1 * R[5] + T2 * R[8]) > ebz + eax * f02 + eay * f12 + eaz * f22;
    if (sep) return false;
    // edge x edge
    sep |= fabsf(T2 * R[3] - T1 * R[6]) > eay * f20 + eaz * f10 + eby * f02 + ebz * f01;
    sep |= fabsf(T2 * R[4] - T1 * R[7]) > eay * f21 + eaz * f11 + ebx * f02 + ebz * f00;
    sep |= fabsf(T2 * R[5] - T1 * R[8]) > eay * f22 + eaz * f12 + ebx * f01 + eby * f00;
    sep |= fabsf(T0 * R[6] - T2 * R[0]) > eax * f20 + eaz * f00 + eby * f12 + ebz * f11;
    sep |= fabsf(T0 * R[7] - T2 * R[1]) > eax * f21 + eaz * f01 + ebx * f12 + ebz * f10;
    sep |= fabsf(T0 * R[8] - T2 * R[2]) > eax * f22 + eaz * f02 + ebx * f11 + eby * f10;
    sep |= fabsf(T1 * R[0] - T0 * R[3]) > eax * f10 + eay * f00 + eby * f22 + ebz * f21;
    sep |= fabsf(T1 * R[1] - T0 * R[4]) > eax * f11 + eay * f01 + ebx * f22 + ebz * f20;
    sep |= fabsf(T1 * R[2] - T0 * R[5]) > eax * f12 + eay * f02 + ebx * f21 + eby * f20;
    return !sep;
}

// The same test for TWO boxes A0, A1 (identity frame: the two children of the node being split) against ONE box B
// placed by (R, t).  Straight-line code (no early outs: some lane of the warp always needs the late axes), with
// everything that depends on B alone computed once.  kEdgeAxes = false drops the 9 edge x edge axes: still
// conservative (a box pair may only be reported overlapping more often), cheaper per test, a few more tests.
constexpr bool kEdgeAxes = CC_EDGE_AXES;
__device__ __forceinline__ void box_overlap_pair(const float* __restrict__ R, const float* __restrict__ t,
                                                 float c0x, float c0y, float c0z, float e0x, float e0y, float e0z,
                                                 float c1x, float c1y, float c1z, float e1x, float e1y, float e1z,
                                                 float cbx, float cby, float cbz, float ebx, float eby, float ebz,
                                                 bool& ov0, bool& ov1)
{
    const float reps = 1e-6f;
    const float Bx = fmaf(R[0], cbx, fmaf(R[1], cby, fmaf(R[2], cbz, t[0])));
    const float By = fmaf(R[3], cbx, fmaf(R[4], cby, fmaf(R[5], cbz, t[1])));
    const float Bz = fmaf(R[6], cbx, fmaf(R[7], cby, fmaf(R[8], cbz, t[2])));
    const float f00 = fabsf(R[0]) + reps, f01 = fabsf(R[1]) + reps, f02 = fabsf(R[2]) + reps;
    const float f10 = fabsf(R[3]) + reps, f11 = fabsf(R[4]) + reps, f12 = fabsf(R[5]) + reps;
    const float f20 = fabsf(R[6]) + reps, f21 = fabsf(R[7]) + reps, f22 = fabsf(R[8]) + reps;
    const float rb0 = ebx * f00 + eby * f01 + ebz * f02;      // B's radius along A's axes
    const float rb1 = ebx * f10 + eby * f11 + ebz * f12;
    const float rb2 = ebx * f20 + eby * f21 + ebz * f22;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const float eax = ch ? e1x : e0x, eay = ch ? e1y : e0y, eaz = ch ? e1z : e0z;
        const float T0 = Bx - (ch ? c1x : c0x), T1 = By - (ch ? c1y : c0y), T2 = Bz - (ch ? c1z : c0z);
        bool sep = fabsf(T0) > eax + rb0;
        sep |= fabsf(T1) > eay + rb1;
        sep |= fabsf(T2) > eaz + rb2;
        sep |= fabsf(T0 * R[0] + T1 * R[3] + T2 * R[6]) > ebx + eax * f00 + eay * f10 + eaz * f20;
        sep |= fabsf(T0 * R[1] + T1 * R[4] + T2 * R[7]) > eby + eax * f01 + eay * f11 + eaz * f21;
        sep |= fabsf(T0 * R[2] + T1 * R[5] + T2 * R[8]) > ebz + eax * f02 + eay * f12 + eaz * f22;
        if (kEdgeAxes) {
            sep |= fabsf(T2 * R[3] - T1 * R[6]) > eay * f20 + eaz * f10 + eby * f02 + ebz * f01;
            sep |= fabsf(T2 * R[4] - T1 * R[7]) > eay * f21 + eaz * f11 + ebx * f02 + ebz * f00;
            sep |= fabsf(T2 * R[5] - T1 * R[8]) > eay * f22 + eaz * f12 + ebx * f01 + eby * f00;
            sep |= fabsf(T0 * R[6] - T2 * R[0]) > eax * f20 + eaz * f00 + eby * f12 + ebz * f11;
            sep |= fabsf(T0 * R[7] - T2 * R[1]) > eax * f21 + eaz * f01 + ebx * f12 + ebz * f10;
            sep |= fabsf(T0 * R[8] - T2 * R[2]) > eax * f22 + eaz * f02 + ebx * f11 + eby * f10;
            sep |= fabsf(T1 * R[0] - T0 * R[3]) > eax * f10 + eay * f00 + eby * f22 + ebz * f21;
            sep |= fabsf(T1 * R[1] - T0 * R[4]) > eax * f11 + eay * f01 + ebx * f22 + ebz * f20;
            sep |= fabsf(T1 * R[2] - T0 * R[5]) > eax * f12 + eay * f02 + ebx * f21 + eby * f20;
        }
        if (ch) ov1 = !sep; else ov0 = !sep;
    }
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

// TriContact (pqp/PQP.cpp:375-459): 17 axes n1, m1, e_i x f_j, e_i x n1, f_i x m1.  The verdict is the AND over all axes,
// so the test is cut in two: the two face normals (which separate most non-touching leaf pairs) run on every queued pair,
// the other 15 axes only on the survivors, re-batched 32 at a time so the expensive part runs on full warps.
// p1, p2, q0, q1, q2 are the vertices after the shift by -P0 (P0 is the origin).
__device__ __forceinline__ bool tri_contact_normals(V3 p1, V3 p2, V3 q0, V3 q1, V3 q2, float eps2)
{
    V3 n1 = cross(p1, p2 - p1);
    if (!axis_overlaps(n1, p1, p2, q0, q1, q2, eps2)) return false;
    V3 m1 = cross(q1 - q0, q2 - q1);
    return axis_overlaps(m1, p1, p2, q0, q1, q2, eps2);
}

__device__ __forceinline__ bool tri_contact_edges(V3 p1, V3 p2, V3 q0, V3 q1, V3 q2, float eps2)
{
    V3 e1 = p1, e2 = p2 - p1, e3 = v3(-p2.x, -p2.y, -p2.z);
    V3 f1 = q1 - q0, f2 = q2 - q1, f3 = q0 - q2;
    V3 n1 = cross(e1, e2);
    V3 m1 = cross(f1, f2);
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

// Vertices of one queued leaf pair in tree A's frame, shifted by -P0.
__device__ __forceinline__ void load_tri_pair(const float4* __restrict__ tris, const float* __restrict__ rt, uint32_t ta_idx, uint32_t tb_idx,
                                              V3& p1, V3& p2, V3& q0, V3& q1, V3& q2)
{
    const float4* ta = tris + 3 * (size_t)ta_idx;
    const float4* tb = tris + 3 * (size_t)tb_idx;
    const float4 a0 = __ldg(ta), a1 = __ldg(ta + 1), a2 = __ldg(ta + 2);
    const float4 u0 = __ldg(tb), u1 = __ldg(tb + 1), u2 = __ldg(tb + 2);
    const float r0 = rt[0], r1 = rt[1], r2 = rt[2], r3 = rt[3], r4 = rt[4], r5 = rt[5], r6 = rt[6], r7 = rt[7], r8 = rt[8];
    const float t0 = rt[9] - a0.x, t1 = rt[10] - a0.y, t2 = rt[11] - a0.z;
    p1 = v3(a1.x - a0.x, a1.y - a0.y, a1.z - a0.z);
    p2 = v3(a2.x - a0.x, a2.y - a0.y, a2.z - a0.z);
    q0 = v3(fmaf(r0, u0.x, fmaf(r1, u0.y, fmaf(r2, u0.z, t0))), fmaf(r3, u0.x, fmaf(r4, u0.y, fmaf(r5, u0.z, t1))),
            fmaf(r6, u0.x, fmaf(r7, u0.y, fmaf(r8, u0.z, t2))));
    q1 = v3(fmaf(r0, u1.x, fmaf(r1, u1.y, fmaf(r2, u1.z, t0))), fmaf(r3, u1.x, fmaf(r4, u1.y, fmaf(r5, u1.z, t1))),
            fmaf(r6, u1.x, fmaf(r7, u1.y, fmaf(r8, u1.z, t2))));
    q2 = v3(fmaf(r0, u2.x, fmaf(r1, u2.y, fmaf(r2, u2.z, t0))), fmaf(r3, u2.x, fmaf(r4, u2.y, fmaf(r5, u2.z, t1))),
            fmaf(r6, u2.x, fmaf(r7, u2.y, fmaf(r8, u2.z, t2))));
}

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

// ================================================================================================ kernel 1: FK
__host__ __device__ inline size_t fk_smem_bytes(int nq, int nsteps, int ndyn)
{
    return (size_t)nsteps * sizeof(FkStep) + (size_t)ndyn * 32 + (size_t)nq * 4;
}

// Where the FK of ONE configuration puts its results (row c of the per-chunk pose scratch).
struct FkDst {
    float* xf;  int xf_stride;       // FP32 R|t of slot d at xf + d * xf_stride (12 floats)
    float* ab;  int ab_stride;       // world AABB of slot d at ab + d * ab_stride (8 floats)
    double* save;                    // parked pose i, component k at save[i * save_pose_stride + k * save_comp_stride]
    size_t save_pose_stride, save_comp_stride;
};

// Emit one link pose: optional FP64 copy-out, FP32 R|t (matrixFromQuat, geometry.h:779-800) and the world AABB
// of the link's root box, both into the dynamic slot `d` of configuration `c` (d < 0: only the copy-out).
__device__ __forceinline__ void emit_link_box(const Params& P, const FkDst& D, int c, int l, int d, const Xf& T,
                                              float cx, float cy, float cz, float ex, float ey, float ez)
{
    if (P.fk_out) {
        double* dst = P.fk_out + ((size_t)c * P.nq + l) * 8;
        dst[0] = T.q0; dst[1] = T.q1; dst[2] = T.q2; dst[3] = T.q3; dst[4] = T.tx; dst[5] = T.ty; dst[6] = T.tz; dst[7] = 0.0;
    }
    if (d < 0 || (P.flags & kFlagFkOnly)) return;
    const double il = 2.0 / (T.q0 * T.q0 + T.q1 * T.q1 + T.q2 * T.q2 + T.q3 * T.q3);
    const double qq1 = il * T.q1 * T.q1, qq2 = il * T.q2 * T.q2, qq3 = il * T.q3 * T.q3;
    float R[9], t[3];
    R[0] = (float)(1 - qq2 - qq3);                      R[1] = (float)(il * (T.q1 * T.q2 - T.q0 * T.q3)); R[2] = (float)(il * (T.q1 * T.q3 + T.q0 * T.q2));
    R[3] = (float)(il * (T.q1 * T.q2 + T.q0 * T.q3));   R[4] = (float)(1 - qq1 - qq3);                    R[5] = (float)(il * (T.q2 * T.q3 - T.q0 * T.q1));
    R[6] = (float)(il * (T.q1 * T.q3 - T.q0 * T.q2));   R[7] = (float)(il * (T.q2 * T.q3 + T.q0 * T.q1)); R[8] = (float)(1 - qq1 - qq2);
    t[0] = (float)(T.tx - P.center[0]); t[1] = (float)(T.ty - P.center[1]); t[2] = (float)(T.tz - P.center[2]);
    float4* dst = (float4*)(D.xf + (size_t)d * D.xf_stride);
    dst[0] = make_float4(R[0], R[1], R[2], R[3]);
    dst[1] = make_float4(R[4], R[5], R[6], R[7]);
    dst[2] = make_float4(R[8], t[0], t[1], t[2]);
    const float reps = 1e-6f;
    float4* ab = (float4*)(D.ab + (size_t)d * D.ab_stride);
    ab[0] = make_float4(fmaf(R[0], cx, fmaf(R[1], cy, fmaf(R[2], cz, t[0]))), fmaf(R[3], cx, fmaf(R[4], cy, fmaf(R[5], cz, t[1]))),
                        fmaf(R[6], cx, fmaf(R[7], cy, fmaf(R[8], cz, t[2]))), 0.f);
    ab[1] = make_float4((fabsf(R[0]) + reps) * ex + (fabsf(R[1]) + reps) * ey + (fabsf(R[2]) + reps) * ez,
                        (fabsf(R[3]) + reps) * ex + (fabsf(R[4]) + reps) * ey + (fabsf(R[5]) + reps) * ez,
                        (fabsf(R[6]) + reps) * ex + (fabsf(R[7]) + reps) * ey + (fabsf(R[8]) + reps) * ez, 0.f);
}

// (link boxes as an array of {cx cy cz ex ey ez . .} records)
__device__ __forceinline__ void emit_link(const Params& P, const float* linkbox, const FkDst& D, int c, int l, int d, const Xf& T)
{
    const float* lb = linkbox + 8 * (d < 0 ? 0 : d);
    emit_link_box(P, D, c, l, d, T, lb[0], lb[1], lb[2], lb[3], lb[4], lb[5]);
}

// L * J(value) * R of one joint (kinbody.cpp:2404-2478); the value rules follow kinbody.cpp:2149-2224, 2308-2386.
__device__ __forceinline__ Xf joint_transform(const FkStep& st, const double* __restrict__ qrow)
{
    Xf L = xf_load(st.L);
    if (st.flags & 1) return L;                      // static joint: right already folded (kinbodyjoint.cpp:1527-1551)
    double v = st.src_col >= 0 ? qrow[st.src_col] : st.src_const;
    if (!(st.flags & 8)) {                           // SetDOFValues clamps non-circular DOFs
        if (v < st.src_lower) v = st.src_lower; else if (v > st.src_upper) v = st.src_upper;
    }
    if (st.flags & 2) {                              // mimic
        const double ev = st.a * v + st.b;
        v = ev;
        if (!(st.flags & 4)) {
            // kinbody.cpp:2326-2375: within g_fEpsilonJointLimit of a limit -> the limit; farther out than g_fEpsilonEvalJointLimit
            // -> clamped (CLA_CheckLimits); the sliver in between keeps the evaluated value, as in the reference
            if (ev < st.lower) { if (ev >= st.lower - CC_EPSILON_JOINT_LIMIT) v = st.lower; else if (ev < st.lower - CC_EPSILON_EVAL_JOINT_LIMIT) v = st.lower; }
            else if (ev > st.upper) { if (ev <= st.upper + CC_EPSILON_JOINT_LIMIT) v = st.upper; else if (ev > st.upper + CC_EPSILON_EVAL_JOINT_LIMIT) v = st.upper; }
        }
    }
    Xf J; J.q0 = 1; J.q1 = J.q2 = J.q3 = 0; J.tx = J.ty = J.tz = 0;
    if (st.type == 1) {
        // quatFromAxisAngle (geometry.h:673-682)
        const double len = sqrt(st.axis[0] * st.axis[0] + st.axis[1] * st.axis[1] + st.axis[2] * st.axis[2]);
        if (len != 0) {
            double sn, cs; sincos(v * 0.5, &sn, &cs);
            sn /= len;
            J.q0 = cs; J.q1 = st.axis[0] * sn; J.q2 = st.axis[1] * sn; J.q3 = st.axis[2] * sn;
        }
    } else {
        J.tx = st.axis[0] * v; J.ty = st.axis[1] * v; J.tz = st.axis[2] * v;
    }
    return xf_mul(xf_mul(L, J), xf_load(st.R));       // (L * J) * R, kinbody.cpp:2478
}

// Link 0's pose of one configuration when the active DOFs include affine ones.  Same bit values as OpenRAVE::DOFAffine.
__device__ __forceinline__ Xf affine_base(const Params& P, Xf t, const double* __restrict__ a)
{
    const int m = P.affine_mask;
    if (m & 1) t.tx = *a++;
    if (m & 2) t.ty = *a++;
    if (m & 4) t.tz = *a++;
    if (m & 8) {                                   // DOF_RotationAxis
        double sn, cs; sincos(*a * 0.5, &sn, &cs);
        t.q0 = cs; t.q1 = P.affine_axis[0] * sn; t.q2 = P.affine_axis[1] * sn; t.q3 = P.affine_axis[2] * sn;
    } else if (m & 16) {                           // DOF_Rotation3D: axis * angle
        const double x = a[0], y = a[1], z = a[2];
        const double ang = sqrt(x * x + y * y + z * z);
        if (ang > 0) {
            double sn, cs; sincos(0.5 * ang, &sn, &cs);
            const double f = sn / ang;
            t.q0 = cs; t.q1 = f * x; t.q2 = f * y; t.q3 = f * z;
        } else { t.q0 = 1; t.q1 = t.q2 = t.q3 = 0; }
    } else if (m & 32) {                           // DOF_RotationQuat: normalize4, then _vRotationQuatLimitStart * q (robot.cpp:1205-1207)
        double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3];
        quat_normalize(r0, r1, r2, r3);
        const double* s = P.affine_qstart;
        t.q0 = s[0] * r0 - s[1] * r1 - s[2] * r2 - s[3] * r3;
        t.q1 = s[0] * r1 + s[1] * r0 + s[2] * r3 - s[3] * r2;
        t.q2 = s[0] * r2 + s[2] * r0 + s[3] * r1 - s[1] * r3;
        t.q3 = s[0] * r3 + s[3] * r0 + s[1] * r2 - s[2] * r1;
    }
    return t;
}

// KinBody::SetDOFValues' joint loop for configuration c (kinbody.cpp:2286-2481): the chain program is walked depth-first with
// the running transform in registers; poses later branches start from are parked through D.save.
__device__ __forceinline__ void fk_chain(const Params& P, const FkStep* steps, const float* linkbox, const int32_t* emit, const FkDst& D, int c)
{
    const double* qrow = P.q + (size_t)c * P.ncols;
    Xf base = xf_load(P.base);                                    // link 0 keeps its pose (kinbody.cpp:2281) ...
    if (P.affine_mask) base = affine_base(P, base, qrow + P.affine_col);      // ... unless affine DOFs are active
    emit_link(P, linkbox, D, c, 0, emit[0], base);
    Xf cur = base;
    const size_t ps = D.save_pose_stride, cs = D.save_comp_stride;
    for (int s = 0; s < P.nsteps; ++s) {
        const FkStep& st = steps[s];
        if (st.load_parent == -1) cur = base;
        else if (st.load_parent >= 0) {
            const double* sv = D.save + (size_t)st.load_parent * ps;
            cur.q0 = sv[0]; cur.q1 = sv[cs]; cur.q2 = sv[2 * cs]; cur.q3 = sv[3 * cs];
            cur.tx = sv[4 * cs]; cur.ty = sv[5 * cs]; cur.tz = sv[6 * cs];
        }
        if (st.flags & 16) {                                     // rigidly attached to the running link: pose out, nothing else
            emit_link(P, linkbox, D, c, st.child, emit[st.child], xf_mul(cur, xf_load(st.L)));
            continue;
        }
        cur = xf_mul(cur, joint_transform(st, qrow));
        if (st.save >= 0) {
            double* sv = D.save + (size_t)st.save * ps;
            sv[0] = cur.q0; sv[cs] = cur.q1; sv[2 * cs] = cur.q2; sv[3 * cs] = cur.q3;
            sv[4 * cs] = cur.tx; sv[5 * cs] = cur.ty; sv[6 * cs] = cur.tz;
        }
        emit_link(P, linkbox, D, c, st.child, emit[st.child], cur);
    }
}

#ifndef CC_FK_BLOCKS
#define CC_FK_BLOCKS 4      // 104 registers, no spilled loop counter: 4 > 5 > 6 blocks by 0.5 - 2 % (profiles/r02l_ab.log)
#endif
__global__ void __launch_bounds__(kFkThreads, CC_FK_BLOCKS)
cc_fk_kernel(const __grid_constant__ Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FkStep* s_steps = (FkStep*)smem_raw;
    float* s_linkbox = (float*)(s_steps + P.nsteps);
    int32_t* s_emit = (int32_t*)(s_linkbox + 8 * P.ndyn);
    {
        const int4* src = (const int4*)P.steps; int4* dst = (int4*)s_steps;
        for (int i = threadIdx.x; i < P.nsteps * (int)(sizeof(FkStep) / 16); i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < 8 * P.ndyn; i += blockDim.x) s_linkbox[i] = P.link_box[i];
        for (int i = threadIdx.x; i < P.nq; i += blockDim.x) s_emit[i] = P.emit_index[i];
    }
    __syncthreads();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n) return;

    if (P.flags & kFlagInitOutputs) {
        if (P.verdict) P.verdict[c] = 0u;
        if (P.first_env) P.first_env[c] = -1;
        if (P.first_self) P.first_self[c] = -1;
        if (P.work_out) P.work_out[c] = 0u;
        if (P.pair_mask) {
            uint32_t* mrow = P.pair_mask + (size_t)c * P.mask_words;
            for (int i = 0; i < P.mask_words; ++i) mrow[i] = P.hoisted_mask ? P.hoisted_mask[i] : 0u;
        }
    }
    if (P.flags & kFlagPoseMode) {
        const double* src = P.q + (size_t)c * P.nq * 8;
        for (int l = 0; l < P.nq; ++l)
            if (s_emit[l] >= 0) {
                FkDst D;
                D.xf = P.xf32 + (size_t)c * P.ndyn * 12; D.xf_stride = 12;
                D.ab = P.aabb + (size_t)c * P.ndyn * 8; D.ab_stride = 8;
                D.save = nullptr; D.save_pose_stride = D.save_comp_stride = 0;
                emit_link(P, s_linkbox, D, c, l, s_emit[l], xf_load(src + 8 * l));
            }
        return;
    }
    FkDst D;
    D.xf = P.xf32 + (size_t)c * P.ndyn * 12; D.xf_stride = 12;
    D.ab = P.aabb + (size_t)c * P.ndyn * 8; D.ab_stride = 8;
    D.save = P.fksave + c; D.save_pose_stride = 7 * (size_t)P.cap; D.save_comp_stride = (size_t)P.cap;
    fk_chain(P, s_steps, s_linkbox, s_emit, D, c);
}

// ------------------------------------------------------------------------------------------------ kernel 1, two-phase form
// The same arithmetic as cc_fk_kernel, split so that most of it runs with every thread busy and with coalesced stores:
//   phase 1  one thread per configuration walks only the VARYING steps (8 of PR2's 41) and leaves every varying link's pose in
//            shared memory ([configuration][pose][7 doubles], rows padded to an odd length: conflict-free both ways);
//   phase 2  one thread per (configuration, link): a varying link's pose is read back, a rigidly attached link's pose is
//            T_anchor * L (one product), then the FP32 R|t, the world AABB and the optional FP64 copy-out.  Consecutive threads
//            write consecutive slots of one configuration's row, where the one-thread-per-configuration kernel scatters every
//            16-byte store over 32 rows and walks all 41 links behind one serial dependency chain (issue-active 18 %,
//            long-scoreboard 14.4, profiles/r02i_fk_summary.md).
__host__ __device__ inline size_t fk2_fixed_bytes(int nq, int nsteps, int ndyn)
{
    // chain program | link -> slot | per item of phase 2 (link 0 + one per step), structure-of-arrays: 7 doubles of L, then
    // {flags, pose_in, pose_out, child} | link boxes, structure-of-arrays (6 floats per dynamic slot)
    return ((size_t)nsteps * sizeof(FkStep) + (size_t)nq * 4 + 15) / 16 * 16 + ((size_t)(nsteps + 1) * (56 + 16) + (size_t)ndyn * 24 + 15) / 16 * 16;
}
__host__ __device__ inline int fk2_row_doubles(int npose) { return (npose * 7) | 1; }

__device__ __forceinline__ Xf pose_load(const double* row, int idx)
{
    const double* p = row + 7 * idx;
    Xf t; t.q0 = p[0]; t.q1 = p[1]; t.q2 = p[2]; t.q3 = p[3]; t.tx = p[4]; t.ty = p[5]; t.tz = p[6];
    return t;
}
__device__ __forceinline__ void pose_store(double* row, int idx, const Xf& t)
{
    double* p = row + 7 * idx;
    p[0] = t.q0; p[1] = t.q1; p[2] = t.q2; p[3] = t.q3; p[4] = t.tx; p[5] = t.ty; p[6] = t.tz;
}

// grid: one block per `group` configurations (group <= blockDim.x); dynamic shared memory: fk2_fixed_bytes + group * row * 8
__global__ void __launch_bounds__(kFkThreads, 4)
cc_fk2_kernel(const __grid_constant__ Params P, const int group)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FkStep* s_steps = (FkStep*)smem_raw;
    int32_t* s_emit = (int32_t*)(s_steps + P.nsteps);
    const int ni = P.nsteps + 1;                                    // items of phase 2 per configuration: link 0 + one per step
    double* s_L = (double*)(smem_raw + ((size_t)P.nsteps * sizeof(FkStep) + (size_t)P.nq * 4 + 15) / 16 * 16);      // [7][ni]
    int32_t* s_meta = (int32_t*)(s_L + 7 * ni);                     // [4][ni]: flags, pose_in, pose_out, child
    float* s_lb = (float*)(s_meta + 4 * ni);                        // [6][ndyn]
    double* s_pose = (double*)(smem_raw + fk2_fixed_bytes(P.nq, P.nsteps, P.ndyn));
    const int rowlen = fk2_row_doubles(P.npose);
    {
        const int4* src = (const int4*)P.steps; int4* dst = (int4*)s_steps;
        for (int i = threadIdx.x; i < P.nsteps * (int)(sizeof(FkStep) / 16); i += blockDim.x) dst[i] = src[i];
        for (int i = threadIdx.x; i < P.nq; i += blockDim.x) s_emit[i] = P.emit_index[i];
        for (int i = threadIdx.x; i < 6 * P.ndyn; i += blockDim.x) { const int k = i / P.ndyn, d = i - k * P.ndyn; s_lb[i] = P.link_box[8 * d + k]; }
    }
    __syncthreads();
    // phase 2 reads its step constants with consecutive threads on consecutive steps: as 256-byte records every such load is a
    // 32-way bank conflict (l1tex 87 % of peak in profiles/r02x_fk2_summary.md, 49 % after in r02y_fk2_summary.md), as arrays it is conflict-free
    for (int e = threadIdx.x; e < ni; e += blockDim.x) {
        if (e == 0) {
            s_meta[0] = 0; s_meta[ni] = 0; s_meta[2 * ni] = 0; s_meta[3 * ni] = 0;
            for (int k = 0; k < 7; ++k) s_L[k * ni] = k == 0 ? 1.0 : 0.0;
        } else {
            const FkStep& st = s_steps[e - 1];
            s_meta[e] = st.flags; s_meta[ni + e] = st.pose_in; s_meta[2 * ni + e] = st.pose_out; s_meta[3 * ni + e] = st.child;
            for (int k = 0; k < 7; ++k) s_L[k * ni + e] = st.L[k];
        }
    }
    const int c0 = blockIdx.x * group;
    const int nc = P.n - c0 < group ? P.n - c0 : group;
    const bool pose_mode = (P.flags & kFlagPoseMode) != 0;
    // ---- phase 1: the varying steps, one thread per configuration
    for (int i = threadIdx.x; i < nc; i += blockDim.x) {
        const int c = c0 + i;
        if (P.flags & kFlagInitOutputs) {
            if (P.verdict) P.verdict[c] = 0u;
            if (P.first_env) P.first_env[c] = -1;
            if (P.first_self) P.first_self[c] = -1;
            if (P.work_out) P.work_out[c] = 0u;
            if (P.pair_mask) {
                uint32_t* mrow = P.pair_mask + (size_t)c * P.mask_words;
                for (int k = 0; k < P.mask_words; ++k) mrow[k] = P.hoisted_mask ? P.hoisted_mask[k] : 0u;
            }
        }
        if (pose_mode) continue;
        const double* qrow = P.q + (size_t)c * P.ncols;
        double* row = s_pose + (size_t)i * rowlen;
        Xf base = xf_load(P.base);                                    // link 0 keeps its pose (kinbody.cpp:2281) ...
        if (P.affine_mask) base = affine_base(P, base, qrow + P.affine_col);      // ... unless affine DOFs are active
        pose_store(row, 0, base);
        for (int s = 0; s < P.nsteps; ++s) {
            const FkStep& st = s_steps[s];
            if (st.flags & 16) continue;
            pose_store(row, st.pose_out, xf_mul(pose_load(row, st.pose_in), joint_transform(st, qrow)));
        }
    }
    __syncthreads();
    // ---- phase 2: one thread per (configuration, link)
    const int per_cfg = pose_mode ? P.nq : P.nsteps + 1;
    for (int it = threadIdx.x; it < nc * per_cfg; it += blockDim.x) {
        const int i = it / per_cfg, e = it - i * per_cfg;
        const int c = c0 + i;
        FkDst D;
        D.xf = P.xf32 + (size_t)c * P.ndyn * 12; D.xf_stride = 12;
        D.ab = P.aabb + (size_t)c * P.ndyn * 8; D.ab_stride = 8;
        D.save = nullptr; D.save_pose_stride = D.save_comp_stride = 0;
        int l, d;
        Xf T;
        if (pose_mode) {
            l = e; d = s_emit[e];
            if (d < 0) continue;
            T = xf_load(P.q + ((size_t)c * P.nq + e) * 8);
        } else {
            const double* row = s_pose + (size_t)i * rowlen;
            const int flags = s_meta[e];
            l = s_meta[3 * ni + e]; d = s_emit[l];
            if (flags & 16) {
                Xf L; L.q0 = s_L[e]; L.q1 = s_L[ni + e]; L.q2 = s_L[2 * ni + e]; L.q3 = s_L[3 * ni + e];
                L.tx = s_L[4 * ni + e]; L.ty = s_L[5 * ni + e]; L.tz = s_L[6 * ni + e];
                T = xf_mul(pose_load(row, s_meta[ni + e]), L);
            } else T = pose_load(row, s_meta[2 * ni + e]);          // (item 0: link 0, pose 0)
        }
        const int dd = d < 0 ? 0 : d;
        emit_link_box(P, D, c, l, d, T, s_lb[dd], s_lb[P.ndyn + dd], s_lb[2 * P.ndyn + dd], s_lb[3 * P.ndyn + dd], s_lb[4 * P.ndyn + dd], s_lb[5 * P.ndyn + dd]);
    }
}

// ================================================================================================ kernel 2: traversal
constexpr size_t kWarpSmemBytes = (size_t)kNumSlots * 96 + (size_t)kNumSlots * 4 + (size_t)kStackCap * 8 + (size_t)kTriQCap * 8 +
                                  (size_t)kTriQ2Cap * 8 + (size_t)kCandCap * 4;

// Warp-cooperative in-place compaction of a queue of {x | slot << 26, y} entries: drops the entries of slots that are dead
// (their pair already has a contact, or a pair that precedes them in reference order has one).  Runs only when a contact
// was just found, so the descent loop itself never has to ask whether an item is still wanted.
__device__ __forceinline__ int filter_queue(uint2* q, int count, uint32_t dead_slots, int lane, unsigned lt_mask)
{
    int w = 0;
    for (int b = 0; b < count; b += 32) {
        const int i = b + lane;
        uint2 it = make_uint2(0u, 0u);
        bool keep = false;
        if (i < count) { it = q[i]; keep = !((dead_slots >> (it.x >> 26)) & 1u); }
        const unsigned m = __ballot_sync(kFull, keep);
        __syncwarp();
        if (keep) q[w + __popc(m & lt_mask)] = it;
        w += __popc(m);
        __syncwarp();
    }
    return w;
}

// Is configuration `qrow` inside a certified-free cell of table tb?  The DOF values are clamped like KinBody::SetDOFValues
// clamps them (kinbody.cpp:2149-2224); a circular DOF outside its range is left to the narrowphase.
__device__ __noinline__ bool table_cell_free(const Params& P, int tb, const double* __restrict__ qrow)
{
    const PairTable& T = P.tables[tb];
    const double two_pi = 6.283185307179586476925286766559;
    double v0 = qrow[T.col0];
    if (v0 < T.lo0 || v0 > T.hi0) {
        if (!(v0 == v0) || T.circ0 == 1) return false;
        if (T.circ0 == 2) { v0 -= two_pi * floor((v0 - T.lo0) / two_pi); if (v0 < T.lo0 || v0 > T.hi0) return false; }
        else v0 = v0 < T.lo0 ? T.lo0 : T.hi0;
    }
    int i0 = (int)((v0 - T.lo0) * T.inv0);
    if (i0 >= T.n0) i0 = T.n0 - 1;
    uint32_t bit = (uint32_t)i0;
    if (T.col1 >= 0) {
        double v1 = qrow[T.col1];
        if (v1 < T.lo1 || v1 > T.hi1) {
            if (!(v1 == v1) || T.circ1 == 1) return false;
            if (T.circ1 == 2) { v1 -= two_pi * floor((v1 - T.lo1) / two_pi); if (v1 < T.lo1 || v1 > T.hi1) return false; }
            else v1 = v1 < T.lo1 ? T.lo1 : T.hi1;
        }
        int i1 = (int)((v1 - T.lo1) * T.inv1);
        if (i1 >= T.n1) i1 = T.n1 - 1;
        bit = (uint32_t)i0 * (uint32_t)T.n1 + (uint32_t)i1;
    }
    bit += T.bit_base;
    return (__ldg(&P.table_bits[bit >> 5]) >> (bit & 31)) & 1u;
}

// ================================================================================================ kernel 1b: broadphase
// World-AABB test of every (configuration, device pair) of a chunk, plus the certificate lookup where a table covers the pair:
// one bit per pair, one row of cand_words words per configuration.  A thread block takes `group` (<= 32) consecutive
// configurations: their boxes are one contiguous span of the FK kernel's output (1.3 KB per PR2 configuration), read once with
// coalesced 256-bit loads and kept in shared memory as 15-bit fixed-point {min, max} triples, 16 bytes per box (a 32 m cube around
// the query body's base, 0.98 mm steps, min rounded down and max up, saturating: a monotone map, so an overlap is never lost; the
// six comparisons are four packed subtractions).
// Each warp then owns words of 32 pairs (lane = pair): the pair is decoded once, a static slot's box is quantised once and stays
// in registers, a dynamic slot's box is ONE shared-memory load per configuration — pairs are in the reference's order, link by
// link, so the lanes of a word mostly read the same box of slot a (a broadcast).  Lane i keeps configuration i's ballot and the
// word goes out with one store.  History (profiles/): sweeping configurations straight from global memory costs a dependent DRAM
// round trip per step (682 us per 262 144-configuration PR2 chunk, r02n_broad_summary.md); FP32 boxes in shared memory made the
// kernel shared-memory-bound (18 wavefronts per word and configuration, l1tex 96 %, 463 us, r02o_broad_summary.md).  In the
// traversal kernel the same test decodes the pair for every configuration, compacts candidates step by step and holds a
// 72-register warp while it waits for the boxes: 1 450 of 7 430 warp instructions per PR2 configuration and 28 % of its stall
// samples (r02i_trav_sass.txt).
constexpr int kBroadMaxWarps = 16;
constexpr float kBroadHalf = 16.0f;                            // half-size of the quantisation cube (metres, base-centred)
constexpr float kBroadScale = 32767.0f / (2.0f * kBroadHalf);          // 15-bit steps of 0.98 mm

struct Box8 { float4 c, e; };
// one 32-byte box {cx cy cz . | ex ey ez .} with a single 256-bit load (LDG.E.256, sm_100); p is 32-byte aligned
__device__ __forceinline__ Box8 ld_box(const float4* p)
{
    Box8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.c.x), "=f"(r.c.y), "=f"(r.c.z), "=f"(r.c.w), "=f"(r.e.x), "=f"(r.e.y), "=f"(r.e.z), "=f"(r.e.w) : "l"(p));
    return r;
}

// {minx | miny << 16, minz, maxx | maxy << 16, maxz} in 15-bit steps: min rounded down, max rounded up (0.01 of a step of slack
// for the FP32 rounding of the scaling itself), both saturating
__host__ __device__ __forceinline__ uint4 quantise_box(const Box8& b)
{
    auto lo = [](float v) -> uint32_t { return (uint32_t)fminf(fmaxf(floorf((v + kBroadHalf) * kBroadScale - 0.01f), 0.f), 32767.f); };
    auto hi = [](float v) -> uint32_t { return (uint32_t)fminf(fmaxf(ceilf((v + kBroadHalf) * kBroadScale + 0.01f), 0.f), 32767.f); };
    return make_uint4(lo(b.c.x - b.e.x) | (lo(b.c.y - b.e.y) << 16), lo(b.c.z - b.e.z), hi(b.c.x + b.e.x) | (hi(b.c.y + b.e.y) << 16), hi(b.c.z + b.e.z));
}

// a.min <= b.max and b.min <= a.max on all three axes, two 15-bit fields per subtraction: (max | guard) - min keeps the guard bit
// (bit 15 of its half) exactly when max >= min, and never borrows across the halves
__host__ __device__ __forceinline__ bool qbox_overlap(const uint4 a, const uint4 b)
{
    constexpr uint32_t G = 0x80008000u;
    const uint32_t r = ((b.z | G) - a.x) & ((b.w | G) - a.y) & ((a.z | G) - b.x) & ((a.w | G) - b.y);
    return (r & G) == G;
}

// grid: one block per `group` configurations; block: blockDim.x / 32 warps, warp i takes the words i, i + warps, ...;
// dynamic shared memory: group * ndyn * 16 bytes
__global__ void __launch_bounds__(kBroadMaxWarps * 32)
cc_broad_kernel(const __grid_constant__ Params P, const int group)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* s_rows = (uint4*)smem_raw;                                          // [group][ndyn]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int ndyn = P.ndyn;
    const float4* __restrict__ sab = (const float4*)P.st_aabb;
    const int kbeg = (P.flags & kFlagEnv) ? 0 : P.nenv_dev;
    const int kend = (P.flags & kFlagSelf) ? P.npairs_dev : P.nenv_dev;
    const int w_beg = kbeg >> 5, w_end = (kend + 31) >> 5;
    const int c0 = blockIdx.x * group;
    const int nc = P.n - c0 < group ? P.n - c0 : group;
    {
        const float4* __restrict__ src = (const float4*)P.aabb + (size_t)c0 * ndyn * 2;
        for (int i = threadIdx.x; i < nc * ndyn; i += blockDim.x) s_rows[i] = quantise_box(ld_box(src + 2 * i));
    }
    __syncthreads();
    for (int w = w_beg + wib; w < w_end; w += nwarps) {
        const int k = 32 * w + lane;
        const bool live = k < kend;
        const int kk = live ? k : kend - 1;
        const uint32_t pr = __ldg(&P.pairs[kk]);
        const int sa = pr & 0xffff, sb = pr >> 16;
        const bool a_dyn = sa < ndyn, b_dyn = sb < ndyn;
        const uint4* pa = s_rows + (a_dyn ? sa : 0);
        const uint4* pb = s_rows + (b_dyn ? sb : 0);
        uint4 A = make_uint4(0u, 0u, 0u, 0u), B = A;
        if (!a_dyn) A = quantise_box(ld_box(sab + 2 * (sa - ndyn)));
        if (!b_dyn) B = quantise_box(ld_box(sab + 2 * (sb - ndyn)));
        int tb = -1;
        if (P.pair_table && live) tb = __ldg(&P.pair_table[kk]);
        // warp-uniform flavours of the sweep: a word of (moving link, environment link) pairs reads only slot a's box per
        // configuration; a word with a certificate table also looks the configuration's cell up
        const bool env_word = __all_sync(kFull, a_dyn && !b_dyn);
        const bool tables = __any_sync(kFull, tb >= 0);
        uint32_t mine = 0u;
        if (env_word && !tables) {
#pragma unroll 4
            for (int i = 0; i < nc; ++i) {
                A = *pa; pa += ndyn;
                const unsigned m = __ballot_sync(kFull, live & qbox_overlap(A, B));
                if (lane == i) mine = m;
            }
        } else {
#pragma unroll 2
            for (int i = 0; i < nc; ++i) {
                if (a_dyn) A = *pa;
                if (b_dyn) B = *pb;
                pa += ndyn; pb += ndyn;
                bool ov = live & qbox_overlap(A, B);
                if (ov && tb >= 0 && table_cell_free(P, tb, P.q + (size_t)(c0 + i) * P.ncols)) ov = false;
                const unsigned m = __ballot_sync(kFull, ov);
                if (lane == i) mine = m;
            }
        }
        if (lane < nc) P.cand[(size_t)(c0 + lane) * P.cand_words + w] = mine;
    }
}

// kSplit = false: one work item per configuration (full batches).  kSplit = true: P.split work items per configuration,
// each over its own share of the pair list, outputs merged with atomics (small batches: the latency of a batch is the
// traversal of its slowest configuration, so that one is spread over several warps).
// kTables: the plan carries certified free-space tables (compiled out otherwise, so plans without tables run the plain code).
// kContacts: all contacts of every pair are enumerated and recorded (PQP_ALL_CONTACTS, collisionPQP.h:578): no pair is
// abandoned at its first contact; used with kFlagAllPairs for single configurations.
// kCand: the broadphase has already run (cc_broad_kernel, certificates included): candidates come from P.cand, 32 pairs per lane.
template <bool kSplit, bool kTables, bool kContacts = false, bool kCand = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, kMinBlocksPerSM)
cc_traverse_kernel(const __grid_constant__ Params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + wib * kWarpSmemBytes;
    float* slot_rt = (float*)wbase;                               // [kNumSlots][24]: R|t mapping tree B into tree A, then its inverse
    int32_t* slot_pair = (int32_t*)(slot_rt + kNumSlots * 24);    // [kNumSlots] original pair index
    uint2* stack = (uint2*)(slot_pair + kNumSlots);               // [kStackCap] {x | slot<<26, y | side<<31}
    uint2* triq = stack + kStackCap;                              // [kTriQCap]  {triA | slot<<26, triB}
    uint2* triq2 = triq + kTriQCap;                               // [kTriQ2Cap] same, after the face-normal axes
    uint32_t* cbuf = (uint32_t*)(triq2 + kTriQ2Cap);              // [kCandCap]  device pair indices
    uint2* spill = P.spill + (size_t)(blockIdx.x * kWarpsPerBlock + wib) * P.spill_cap;

    const bool all_pairs = (P.flags & kFlagAllPairs) != 0;
    const bool first_hit = (P.flags & kFlagFirstHit) != 0 && !all_pairs;
    const bool counters = (P.flags & kFlagCounters) != 0;
    const float eps2 = P.tri_eps * P.tri_eps;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int nenv_total = P.nenv_total, ndyn = P.ndyn;
    int nenv_dev = P.nenv_dev, npairs_dev = P.npairs_dev;
    const uint32_t* pairs = P.pairs;
    const int32_t* pair_orig = P.pair_orig;
    const int32_t* ptable = P.pair_table;
    const int nwork = kSplit ? P.n * P.split : P.n;
    unsigned long long n_bv = 0, n_tri = 0;
    // pose scratch: the slots of one configuration are contiguous ([cap][ndyn][12] / [cap][ndyn][8]).  Measured and left out in
    // round 2 (A/B logs under profiles/): slot-major scratch (coalesced FK stores, scattered traversal reads: -5 % PR2, -30 % clutter,
    // r02c_ab_fk.log); FK inside this kernel, one warp per 32 configurations (46 vs 81 M configs/s: the latency-bound FK occupies
    // warp residency, r02b_ab_fused_fk.log); staging the configuration's link boxes in shared memory for the broadphase (-4 .. 10 %,
    // r02g_ab_broad_smem.log); giving up on heavy configurations and redoing them split 8 ways (r02e_ab_defer.log); handing LIFO items
    // of heavy configurations to idle warps through global records (the extra state spills at the 72-register cap: 59 / 45 vs
    // 81 M configs/s, r02h_ab_share.log).
    constexpr int xs = 12, as = 8;

    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&P.counters[0], 1);
        c = __shfl_sync(kFull, c, 0);
        if (c >= nwork) break;
        if (kSplit) {
            const int w = c;
            c = w / P.split;
            const int part = w - c * P.split;
            const int pb = __ldg(&P.part_begin[part]);
            pairs = P.pairs + pb; pair_orig = P.pair_orig + pb;
            if (kTables) ptable = P.pair_table + pb;
            npairs_dev = __ldg(&P.part_begin[part + 1]) - pb;
            nenv_dev = __ldg(&P.part_nenv[part]);
        }
        const float* xf = P.xf32 + (size_t)c * ndyn * 12;
        const float* ab = P.aabb + (size_t)c * ndyn * 8;
        uint32_t* mask_row = (all_pairs && P.pair_mask) ? P.pair_mask + (size_t)c * P.mask_words : nullptr;
        if (!kSplit && P.pair_mask) {
            uint32_t* mrow = P.pair_mask + (size_t)c * P.mask_words;
            for (int i = lane; i < P.mask_words; i += 32) mrow[i] = P.hoisted_mask ? P.hoisted_mask[i] : 0u;
            __syncwarp();
        }

        // warp-uniform state, all in registers
        int sp = 0, spill_count = 0, ntri = 0, ntri2 = 0, ncb = 0;
        int best_env = P.hoisted_first_env, best_self = P.hoisted_first_self;
        if (first_hit) {
            // "a hit precedes everything": -1 makes every remaining pair of the class irrelevant
            if (best_env != kNone) { best_env = -1; best_self = -1; }
            else if (best_self != kNone) best_self = -1;
        }
        uint32_t slot_hit = 0;     // slots of the current batch of root pairs with a contact
        uint32_t slot_dead = 0;    // slots whose items are no longer wanted (contact found, or preceded by a contact); non-zero
                                   // also means refills from the spill area have to be filtered
        bool failed = false;
        uint32_t iters = P.hoisted_first_self != kNone ? 0x80000000u : 0u;     // bit 31: a self contact was really found (first-hit mode)
        bool recheck = true;       // a new best may let the broadphase skip the rest of its class
        int k0 = (P.flags & kFlagEnv) ? 0 : nenv_dev;
        const int kend = (P.flags & kFlagSelf) ? npairs_dev : nenv_dev;

        while ((k0 < kend || ncb > 0) && !failed) {
            if (k0 < kend && ncb < 32) {
                if (!all_pairs && recheck) {
                    // pairs are in reference order inside each class: once a hit precedes everything left, skip the rest
                    recheck = false;
                    if (k0 < nenv_dev) { if (best_env < __ldg(&pair_orig[k0])) { k0 = nenv_dev; recheck = true; continue; } }
                    else if (best_self < __ldg(&pair_orig[k0])) { k0 = kend; continue; }
                }
                // ------------------------------------------------------------ broadphase: 32 x kBroadPerLane pairs, world AABBs
                const int lim = (k0 < nenv_dev && nenv_dev < kend) ? nenv_dev : kend;     // do not straddle env | self
                if (kCand) {
                    // one word of 32 precomputed pair bits per lane; as many whole words as the candidate buffer takes
                    const int w0 = k0 >> 5, w = w0 + lane;
                    uint32_t bits = 0u;
                    if (32 * w < lim) {
                        bits = P.cand[(size_t)c * P.cand_words + w];
                        if (w == w0) bits &= 0xffffffffu << (k0 & 31);                   // pairs before k0 are done
                        const int over = 32 * w + 32 - lim;                               // pairs from lim on: next class / none
                        if (over > 0) bits &= 0xffffffffu >> over;
                    }
                    const int cntw = __popc(bits);
                    int incl = cntw;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
                    const bool take = ncb + incl <= kCandCap;                            // non-decreasing: a prefix of the lanes
                    const int nw = __popc(__ballot_sync(kFull, take));                    // >= 1: ncb < 32 here, a word adds <= 32
                    if (take) {
                        int off = ncb + incl - cntw;
                        while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1u; cbuf[off++] = (uint32_t)(32 * w + b); }
                    }
                    ncb += __shfl_sync(kFull, incl, nw - 1);
                    __syncwarp();
                    k0 = 32 * (w0 + nw) < lim ? 32 * (w0 + nw) : lim;
                    if (k0 == nenv_dev) recheck = true;
                    if (ncb < 32 && k0 < kend) continue;
                    if (ncb == 0) continue;
                } else {
                uint32_t pr[kBroadPerLane];
#pragma unroll
                for (int h = 0; h < kBroadPerLane; ++h) {
                    const int k = k0 + 32 * h + lane;
                    pr[h] = k < lim ? __ldg(&pairs[k]) : 0xffffffffu;
                }
                bool ov[kBroadPerLane];
#pragma unroll
                for (int h = 0; h < kBroadPerLane; ++h) {
                    ov[h] = false;
                    if (pr[h] != 0xffffffffu) {
                        const int sa = pr[h] & 0xffff, sb = pr[h] >> 16;
                        const float4* pa = (const float4*)(sa < ndyn ? ab + as * sa : P.st_aabb + 8 * (size_t)(sa - ndyn));
                        const float4* pb = (const float4*)(sb < ndyn ? ab + as * sb : P.st_aabb + 8 * (size_t)(sb - ndyn));
                        const float4 ac = pa[0], ae = pa[1], bc = pb[0], be = pb[1];
                        ov[h] = fabsf(ac.x - bc.x) <= ae.x + be.x && fabsf(ac.y - bc.y) <= ae.y + be.y && fabsf(ac.z - bc.z) <= ae.z + be.z;
                        if (kTables && ov[h]) {
                            const int tb = __ldg(&ptable[k0 + 32 * h + lane]);
                            if (tb >= 0 && table_cell_free(P, tb, P.q + (size_t)c * P.ncols)) ov[h] = false;
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < kBroadPerLane; ++h) {
                    const unsigned om = __ballot_sync(kFull, ov[h]);
                    if (ov[h]) cbuf[ncb + __popc(om & lt_mask)] = (uint32_t)(k0 + 32 * h + lane);
                    ncb += __popc(om);
                }
                __syncwarp();
                k0 = (k0 + 32 * kBroadPerLane < lim) ? k0 + 32 * kBroadPerLane : lim;
                if (k0 == nenv_dev) recheck = true;
                if (ncb < 32 && k0 < kend) continue;
                if (ncb == 0) continue;
                }
            }
            {
                // ---------------------------------------------------------------- a batch of <= 32 root pairs
                const int cnt = ncb < 32 ? ncb : 32;
                bool cand = false;
                float R[9], t[3];
                int orig = 0;
                uint32_t ra = 0, rb = 0;
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
                if (lane < cnt) {
                    const uint32_t k = cbuf[lane];
                    orig = __ldg(&pair_orig[k]);
                    const bool is_env = (int)k < nenv_dev;
                    bool skip = false;
                    if (!all_pairs) skip = is_env ? (orig > best_env) : (orig > best_self);
                    if (!skip) {
                        const uint32_t pr = __ldg(&pairs[k]);
                        const int la = pr & 0xffff, lb = pr >> 16;
                        float A[12], Bx[12];
                        {
                            const float4* e = la < ndyn ? (const float4*)(xf + xs * la) : (const float4*)(P.st_xf + 12 * (size_t)(la - ndyn));
                            const float4 e0 = e[0], e1 = e[1], e2 = e[2];
                            A[0] = e0.x; A[1] = e0.y; A[2] = e0.z; A[3] = e0.w; A[4] = e1.x; A[5] = e1.y; A[6] = e1.z; A[7] = e1.w;
                            A[8] = e2.x; A[9] = e2.y; A[10] = e2.z; A[11] = e2.w;
                        }
                        {
                            const float4* e = lb < ndyn ? (const float4*)(xf + xs * lb) : (const float4*)(P.st_xf + 12 * (size_t)(lb - ndyn));
                            const float4 e0 = e[0], e1 = e[1], e2 = e[2];
                            Bx[0] = e0.x; Bx[1] = e0.y; Bx[2] = e0.z; Bx[3] = e0.w; Bx[4] = e1.x; Bx[5] = e1.y; Bx[6] = e1.z; Bx[7] = e1.w;
                            Bx[8] = e2.x; Bx[9] = e2.y; Bx[10] = e2.z; Bx[11] = e2.w;
                        }
                        relative_pose(A, Bx, R, t);
                        ra = (uint32_t)__ldg(&P.slot_root[la]); rb = (uint32_t)__ldg(&P.slot_root[lb]);
#if CC_LD256
                        { const Box8 na = ld_box(&P.nodes[2 * ra]), nb = ld_box(&P.nodes[2 * rb]); a0 = na.c; a1 = na.e; b0 = nb.c; b1 = nb.e; }
#else
                        a0 = __ldg(&P.nodes[2 * ra]); a1 = __ldg(&P.nodes[2 * ra + 1]);
                        b0 = __ldg(&P.nodes[2 * rb]); b1 = __ldg(&P.nodes[2 * rb + 1]);
#endif
                        if (counters) n_bv++;
                        cand = box_overlap(R, t, a0.x, a0.y, a0.z, a1.x, a1.y, a1.z, b0.x, b0.y, b0.z, b1.x, b1.y, b1.z);
                    }
                }
                // drop the consumed candidates
                {
                    const int rest = ncb - cnt;
                    uint32_t mv[kBroadPerLane];
#pragma unroll
                    for (int h = 0; h < kBroadPerLane; ++h) mv[h] = (32 * h + lane < rest) ? cbuf[cnt + 32 * h + lane] : 0u;
                    __syncwarp();
#pragma unroll
                    for (int h = 0; h < kBroadPerLane; ++h) if (32 * h + lane < rest) cbuf[32 * h + lane] = mv[h];
                    ncb = rest;
                    __syncwarp();
                }
                const unsigned cm = __ballot_sync(kFull, cand);
                if (cm == 0u) continue;
                const int nsl = __popc(cm);
                if (cand) {
                    const int s = __popc(cm & lt_mask);
                    float* rt = slot_rt + 24 * s;
#pragma unroll
                    for (int i = 0; i < 9; ++i) rt[i] = R[i];
                    rt[9] = t[0]; rt[10] = t[1]; rt[11] = t[2];
                    // inverse: x_B = R^T x_A - R^T t (items whose split node lives in tree B use it, so the two
                    // children always play the un-transformed role and every lane runs the same code)
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int j = 0; j < 3; ++j) rt[12 + 3 * i + j] = R[3 * j + i];
                    rt[21] = -(R[0] * t[0] + R[3] * t[1] + R[6] * t[2]);
                    rt[22] = -(R[1] * t[0] + R[4] * t[1] + R[7] * t[2]);
                    rt[23] = -(R[2] * t[0] + R[5] * t[1] + R[8] * t[2]);
                    slot_pair[s] = orig;
                }
                // first items: both leaves -> triangle queue; else split the larger box (PQP.cpp:539-544 compares BV
                // sizes; any rule is verdict-neutral).  Reversed so the lowest pair index is popped first.
                {
                    const int la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
                    const bool leafpair = cand && la < 0 && lb < 0;
                    const bool push = cand && !leafpair;
                    const unsigned tm = __ballot_sync(kFull, leafpair);
                    const unsigned pm = __ballot_sync(kFull, push);
                    const uint32_t s = (uint32_t)__popc(cm & lt_mask);
                    if (leafpair) triq[ntri + __popc(tm & lt_mask)] = make_uint2((uint32_t)(~la) | (s << 26), (uint32_t)(~lb));
                    if (push) {
                        uint2 it;
                        if (lb < 0 || (la >= 0 && a1.w > b1.w)) it = make_uint2((uint32_t)la | (s << 26), rb);
                        else it = make_uint2((uint32_t)lb | (s << 26), ra | 0x80000000u);
                        stack[sp + (__popc(pm) - 1 - __popc(pm & lt_mask))] = it;
                    }
                    ntri += __popc(tm);
                    sp += __popc(pm);
                }
                __syncwarp();

                // ---------------------------------------------------------------- drain: descent + triangle batches
                for (;;) {
                    const bool stack_empty = sp == 0 && spill_count == 0;
                    const bool run_edges = ntri2 >= CC_TRI_T2 || (ntri2 > 0 && ntri == 0 && stack_empty);
                    if (run_edges || ntri >= CC_TRI_T1 || (ntri > 0 && stack_empty)) {
                        // ---- up to 32 queued leaf pairs (newest first): the 15 edge axes of survivors, or the 2 face normals
                        uint2* q = run_edges ? triq2 : triq;
                        const int have = run_edges ? ntri2 : ntri;
                        const int tn = have < 32 ? have : 32;
                        bool pass = false;
                        uint2 e = make_uint2(0u, 0u);
                        if (lane < tn) {
                            e = q[have - 1 - lane];
                            V3 p1, p2, q0, q1, q2;
                            load_tri_pair(P.tris, slot_rt + 24 * (e.x >> 26), e.x & 0x3ffffffu, e.y, p1, p2, q0, q1, q2);
                            if (run_edges) pass = tri_contact_edges(p1, p2, q0, q1, q2, eps2);
                            else { if (counters) n_tri++; pass = tri_contact_normals(p1, p2, q0, q1, q2, eps2); }
                        }
                        const unsigned pmask = __ballot_sync(kFull, pass);
                        if (!run_edges) {
                            ntri -= tn;
                            if (pmask) {
                                __syncwarp();
                                if (pass) triq2[ntri2 + __popc(pmask & lt_mask)] = e;
                                ntri2 += __popc(pmask);
                                __syncwarp();
                            }
                            continue;
                        }
                        ntri2 -= tn;
                        if (pmask) {
                            // contacts: record them, then purge everything that no longer matters from the queues
                            const uint32_t s = e.x >> 26;
                            const int pidx = pass ? slot_pair[s] : kNone;
                            if (kContacts && pass) {
                                const uint32_t at = atomicAdd(P.contact_count, 1u);
                                if ((int)at < P.contact_cap) P.contact_out[at] = make_uint4((uint32_t)c, (uint32_t)pidx, e.x & 0x3ffffffu, e.y);
                            }
                            slot_hit |= __reduce_or_sync(kFull, pass ? (1u << s) : 0u);
                            const int pe = __reduce_min_sync(kFull, (pass && pidx < nenv_total) ? pidx : kNone);
                            const int ps = __reduce_min_sync(kFull, (pass && pidx >= nenv_total) ? pidx : kNone);
                            if (pe < best_env) { best_env = pe; recheck = true; }
                            if (ps < best_self) { best_self = ps; recheck = true; }
                            if (ps != kNone) iters |= 0x80000000u;
                            if (first_hit) {
                                if (best_env != kNone) { best_env = -1; best_self = -1; }
                                else if (best_self != kNone) best_self = -1;
                            }
                            if (!kContacts) slot_dead |= slot_hit;
                            if (!all_pairs) {
                                bool d = false;
                                if (lane < nsl) { const int p = slot_pair[lane]; d = (p < nenv_total) ? (p > best_env) : (p > best_self); }
                                slot_dead |= __ballot_sync(kFull, d);
                            }
                            sp = filter_queue(stack, sp, slot_dead, lane, lt_mask);
                            ntri = filter_queue(triq, ntri, slot_dead, lane, lt_mask);
                            ntri2 = filter_queue(triq2, ntri2, slot_dead, lane, lt_mask);
                        }
                        continue;
                    }
                    if (sp == 0) {
                        if (spill_count == 0) break;
                        // refill from the global spill area
                        const int m = spill_count < kSpillChunk ? spill_count : kSpillChunk;
                        const uint2* g = spill + (spill_count - m);
                        for (int i = lane; i < m; i += 32) stack[i] = g[i];
                        spill_count -= m;
                        sp = m;
                        __syncwarp();
                        if (slot_dead) { sp = filter_queue(stack, sp, slot_dead, lane, lt_mask); if (sp == 0) continue; }
                    }
                    // ---- pop up to 32 items; each is (children x, x+1 of one tree) x (node y of the other)
                    const int n = sp < 32 ? sp : 32;
                    ++iters;
                    const bool valid = lane < n;
                    uint2 it = make_uint2(0u, 0u);
                    if (valid) it = stack[sp - 1 - lane];
                    sp -= n;
                    __syncwarp();
                    uint32_t of = 0u;      // bit 0 / 1: child 0 / 1 goes to the triangle queue, bit 2 / 3: child 0 / 1 is pushed back as an item
                    uint2 o0 = make_uint2(0u, 0u), o1 = make_uint2(0u, 0u);     // outputs for child 0 / child 1 (tri entry or item)
                    if (valid) {
                        const uint32_t sbits = it.x & 0xfc000000u;
                        const uint32_t x = it.x & 0x3ffffffu, y = it.y & 0x7fffffffu;
                        const bool side = (it.y >> 31) != 0u;                  // false: x is in tree A, true: x is in tree B
#if CC_LD256
                        // a node is 32 bytes: one 256-bit load each (3 instead of 6 loads per item, half the L1 tag lookups)
                        const Box8 n0 = ld_box(&P.nodes[2 * x]), n1 = ld_box(&P.nodes[2 * x + 2]), no = ld_box(&P.nodes[2 * y]);
                        const float4 c0a = n0.c, c0b = n0.e, c1a = n1.c, c1b = n1.e, oa = no.c, ob = no.e;
#else
                        const float4 c0a = __ldg(&P.nodes[2 * x]), c0b = __ldg(&P.nodes[2 * x + 1]);
                        const float4 c1a = __ldg(&P.nodes[2 * x + 2]), c1b = __ldg(&P.nodes[2 * x + 3]);
                        const float4 oa = __ldg(&P.nodes[2 * y]), ob = __ldg(&P.nodes[2 * y + 1]);
#endif
                        const float* rt = slot_rt + 24 * (it.x >> 26) + (side ? 12 : 0);      // maps the other node into the children's frame
                        float R[9], t[3];
#pragma unroll
                        for (int k = 0; k < 9; ++k) R[k] = rt[k];
                        t[0] = rt[9]; t[1] = rt[10]; t[2] = rt[11];
                        if (counters) n_bv += 2;
                        const int lo = __float_as_int(oa.w);
                        bool ov0, ov1;
                        box_overlap_pair(R, t, c0a.x, c0a.y, c0a.z, c0b.x, c0b.y, c0b.z, c1a.x, c1a.y, c1a.z, c1b.x, c1b.y, c1b.z,
                                         oa.x, oa.y, oa.z, ob.x, ob.y, ob.z, ov0, ov1);
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            if (ch ? ov1 : ov0) {
                                const int lc = __float_as_int(ch ? c1a.w : c0a.w);
                                const float sc = ch ? c1b.w : c0b.w;            // |e|^2 of the child, of the other node: ob.w
                                uint2 out;
                                const bool tq = lc < 0 && lo < 0;
                                if (tq) {
                                    out = side ? make_uint2((uint32_t)(~lo) | sbits, (uint32_t)(~lc))
                                               : make_uint2((uint32_t)(~lc) | sbits, (uint32_t)(~lo));
                                } else if (lo < 0 || (lc >= 0 && sc > ob.w)) {
                                    out = make_uint2((uint32_t)lc | sbits, it.y);                                     // split the child
                                } else {
                                    out = make_uint2((uint32_t)lo | sbits, (x + ch) | (side ? 0u : 0x80000000u));     // split the other
                                }
                                if (ch) o1 = out; else o0 = out;
                                of |= (tq ? 1u : 4u) << ch;
                            }
                        }
                    }
                    const bool tq0 = (of & 1u) != 0u, tq1 = (of & 2u) != 0u, ps0 = (of & 4u) != 0u, ps1 = (of & 8u) != 0u;
                    const unsigned t0m = __ballot_sync(kFull, tq0), t1m = __ballot_sync(kFull, tq1);
                    const unsigned p0m = __ballot_sync(kFull, ps0), p1m = __ballot_sync(kFull, ps1);
#if CC_PUSH_REVERSED
                    // lane 0 popped the top of the stack — the item that comes first in the reference's pair order (the first
                    // items are pushed lowest pair on top, and this rule keeps it that way): its outputs go back on TOP, so the
                    // descent keeps working on the earliest live pair first.  A contact there kills every later pair's items
                    // before they are expanded, which is what the serial reference gets from checking pair by pair.
                    if (t0m | t1m) {
                        const int tot = __popc(t0m) + __popc(t1m);
                        int off = ntri + tot - 1 - (__popc(t0m & lt_mask) + __popc(t1m & lt_mask));
                        if (tq0) triq[off--] = o0;
                        if (tq1) triq[off] = o1;
                        ntri += tot;
                    }
                    if (p0m | p1m) {
                        const int tot = __popc(p0m) + __popc(p1m);
                        int off = sp + tot - 1 - (__popc(p0m & lt_mask) + __popc(p1m & lt_mask));
                        if (ps0) stack[off--] = o0;
                        if (ps1) stack[off] = o1;
                        sp += tot;
                    }
#else
                    if (t0m | t1m) {
                        int off = ntri + __popc(t0m & lt_mask) + __popc(t1m & lt_mask);
                        if (tq0) triq[off++] = o0;
                        if (tq1) triq[off] = o1;
                        ntri += __popc(t0m) + __popc(t1m);
                    }
                    if (p0m | p1m) {
                        int off = sp + __popc(p0m & lt_mask) + __popc(p1m & lt_mask);
                        if (ps1) stack[off++] = o1;
                        if (ps0) stack[off] = o0;
                        sp += __popc(p0m) + __popc(p1m);
                    }
#endif
                    __syncwarp();
                    constexpr int kPushMax = 64;
                    if (sp > kStackCap - kPushMax) {
                        // move the oldest kSpillChunk items to global memory and slide the rest down
                        if (spill_count + kSpillChunk > P.spill_cap) {
                            if (lane == 0) atomicExch(P.err_flag, 1);
                            failed = true; sp = 0; ntri = 0; ntri2 = 0; spill_count = 0;        // the host reports the error
                            __syncwarp();
                            break;
                        }
                        uint2* g = spill + spill_count;
                        for (int i = lane; i < kSpillChunk; i += 32) g[i] = stack[i];
                        spill_count += kSpillChunk;
                        const int rest = sp - kSpillChunk;
                        __syncwarp();
                        for (int b = 0; b < rest; b += 32) {
                            const int i = b + lane;
                            uint2 v = make_uint2(0u, 0u);
                            if (i < rest) v = stack[kSpillChunk + i];
                            __syncwarp();
                            if (i < rest) stack[i] = v;
                            __syncwarp();
                        }
                        sp = rest;
                    }
                }
                // ---------------------------------------------------------------- publish this batch's hits
                if (mask_row && slot_hit) {
                    if (lane < nsl && ((slot_hit >> lane) & 1u)) {
                        const int p = slot_pair[lane];
                        atomicOr(&mask_row[p >> 5], 1u << (p & 31));
                    }
                }
                slot_hit = 0; slot_dead = 0;
                __syncwarp();
            }
        }

        // ---------------------------------------------------------------- verdict
        if (lane == 0) {
            int be = best_env, bs = best_self;
            if (mask_row) {
                // complete masks: derive the first pairs from the bits
                be = kNone; bs = kNone;
                for (int wd = 0; wd < P.mask_words; ++wd) {
                    uint32_t m = __ldcg(&mask_row[wd]);   // bits were set with L2 atomics: bypass L1
                    while (m) {
                        const int bit = __ffs(m) - 1; m &= m - 1;
                        const int p = wd * 32 + bit;
                        if (p < nenv_total) { if (p < be) be = p; } else { if (p < bs) bs = p; }
                    }
                }
            }
            uint32_t v = 0;
            if ((P.flags & kFlagEnv) && be != kNone) v |= 1u;
            if ((P.flags & kFlagSelf) && (first_hit ? (iters >> 31) != 0u : bs != kNone)) v |= 2u;
            if (first_hit) { be = -1; bs = -1; }      // no first pairs in this mode
            if (kSplit) {
                // merge with the other work items of this configuration (-1 reads as the largest unsigned value)
                if (v && P.verdict) atomicOr(&P.verdict[c], v);
                if ((v & 1u) && P.first_env) atomicMin((unsigned int*)&P.first_env[c], (unsigned int)be);
                if ((v & 2u) && P.first_self) atomicMin((unsigned int*)&P.first_self[c], (unsigned int)bs);
                if (P.work_out) atomicAdd(&P.work_out[c], iters & 0x7fffffffu);
            } else {
                if (P.verdict) P.verdict[c] = v;
                if (P.first_env) P.first_env[c] = (v & 1u) ? be : -1;
                if (P.first_self) P.first_self[c] = (v & 2u) ? bs : -1;
                if (P.work_out) P.work_out[c] = iters & 0x7fffffffu;
            }
        }
        __syncwarp();
    }
    if (counters) {
        for (int o = 16; o > 0; o >>= 1) { n_bv += __shfl_xor_sync(kFull, n_bv, o); n_tri += __shfl_xor_sync(kFull, n_tri, o); }
        if (lane == 0) { atomicAdd(&P.stats[0], n_bv); atomicAdd(&P.stats[1], n_tri); }
    }
}

// ================================================================================================ diagnostics: FP32 FMA peak
// 8 independent FMA chains per thread, 4096 rounds: the denominator of the compute-side roofline figure.
__global__ void __launch_bounds__(256) cc_fma_peak_kernel(float* out, float a, float b)
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace ccdev
