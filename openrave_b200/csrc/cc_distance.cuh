// cc_distance.cuh — minimum separation per configuration (CO_Distance), sm_100a.
//
// Reference path (paths relative to the OpenRAVE tree): with CO_Distance pqprave runs PQP_Distance on every link pair a query
// visits and keeps the smallest result together with the two closest points (plugins/pqprave/collisionPQP.h:622-648;
// PQP_Distance / DistanceRecurse plugins/pqprave/pqp/PQP.cpp:663-945, TriDist plugins/pqprave/pqp/TriDist.cpp).  pqprave calls it
// with rel_err = 200, abs_err = 0.001 (collisionPQP.h:48-49): an approximate search whose answer depends on the RSS tree and
// the visiting order.  With rel_err = abs_err = 0 this kernel computes the EXACT minimum (what PQP_Distance returns with both
// tolerances 0) in FP32: |result - exact| is a few 1e-6 m for metre-sized scenes.  Non-zero tolerances apply PQP's pruning rule
// (pqp/PQP.cpp:717-721) to the running minimum, so the result r obeys exact <= r <= max(exact + abs_err, exact * (1 + rel_err)).
//
// One warp per configuration (persistent, work counter), branch and bound over the dual AABB trees of all link pairs at once:
//   1. every pair's world-AABB gap (from the FK kernel's boxes) gives a lower bound; the pair with the smallest bound is
//      descended first so that the running minimum `best2` is tight before the other pairs are looked at;
//   2. pairs whose bound is below best2 are descended 32 at a time (one relative pose per slot in shared memory);
//   3. an iteration pops up to 32 items — the two children of one node against a node of the other tree — and each lane bounds
//      both box pairs from below (sum of squared gaps along A's three axes and along B's three axes, the larger of the two);
//      a child whose bound reaches best2 is dropped, a leaf pair goes to the triangle queue, otherwise the larger box of the
//      pair is split into the next item (the nearer child is stored last, so it is expanded first); queued leaf pairs run 32
//      at a time through the exact triangle-triangle distance (9 edge-edge + 6 vertex-face candidates, 0 when the 17-axis
//      test reports contact);
//   4. a warp min-reduction keeps best2, the pair and the closest points.
// Boxes are inflated by 1e-5 m at build time, so a lower bound can only be too small: nothing that could hold the minimum
// is pruned by rounding.
#pragma once
#include "cc_device.cuh"

namespace ccdev {

constexpr int kDistStackCap = 1024;                 // items per warp
constexpr int kDistTriQCap = 96;                    // pending leaf pairs (evaluated 32 at a time, so full warps run the triangle code)
constexpr size_t kDistWarpSmemBytes = (size_t)kNumSlots * 96 + (size_t)kNumSlots * 8 + (size_t)kDistStackCap * 8 +
                                      (size_t)kDistTriQCap * 8 + 32 * 4;

struct DistOut {
    double* min_distance;       // [n] metres; dist_max when nothing is closer (1e30 when dist_max <= 0 and there is no pair)
    int32_t* min_pair;          // [n] reference-order pair index attaining it, -1 when none
    double* points;             // [n][6] closest points in world coordinates (on the query link, on the other link), or NULL
    float dist_max;             // search radius, <= 0: unbounded
    float bound;                // > 0: a pair OUTSIDE this launch is known to be this close (the pose-invariant pairs of the query,
                                // measured once): the search starts from it instead of from the radius; <= 0: none
    // PQP_Distance's tolerances (pqp/PQP.cpp:717-721, 776): a box pair is skipped when its bound d satisfies
    // d + abs_err >= best or d * (1 + rel_err) >= best, so the result r obeys exact <= r <= max(exact + abs_err, exact * (1 + rel_err))
    float rel_err, abs_err;
    // per-pair mode (table building): work item w = configuration w / npairs_dev, pair w % npairs_dev, outputs indexed by w
    int32_t per_pair;
};

__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }

// closest points of segments [p, p + d1] and [q, q + d2] (clamped quadratic minimisation).  Branch-free: the 32 lanes of a
// triangle batch take different clamping cases, and the divergent form ran every case one after the other with 8 - 11 lanes
// each; its unrolled code also overflowed the instruction cache (no_instruction was the top stall, profiles/r02n_dist_summary.md)
__device__ __forceinline__ float seg_seg(V3 p, V3 d1, V3 q, V3 d2, V3& c1, V3& c2)
{
    const V3 r = p - q;
    const float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2);
    const float tiny = 1e-20f;
    const float inv_a = a > tiny ? __fdividef(1.f, a) : 0.f;                 // a degenerate edge keeps its parameter at 0
    const float inv_e = e > tiny ? __fdividef(1.f, e) : 0.f;
    const float den = a * e - b * b;
    float s = den > tiny * a * e ? __saturatef(__fdividef(b * f - c * e, den)) : 0.f;
    const float tn = fmaf(b, s, f);                                          // t = tn / e before clamping
    const float s_lo = __saturatef(-c * inv_a), s_hi = __saturatef((b - c) * inv_a);
    s = (tn < 0.f || e <= tiny) ? s_lo : (tn > e ? s_hi : s);               // t clamped to 0 / 1: s re-solved for that end point
    const float t = __saturatef(tn * inv_e);
    c1 = p + d1 * s;
    c2 = q + d2 * t;
    const V3 d = c1 - c2;
    return dot(d, d);
}

// closest point of triangle (a, b, c) to p: the Voronoi-region walk as selects, latest region first (same priority as the
// early-return form: vertex a, vertex b, edge ab, vertex c, edge ac, edge bc, interior)
__device__ __forceinline__ V3 closest_on_triangle(V3 p, V3 a, V3 b, V3 c)
{
    const V3 ab = b - a, ac = c - a, ap = p - a, bp = p - b, cp = p - c;
    const float d1 = dot(ab, ap), d2 = dot(ac, ap), d3 = dot(ab, bp), d4 = dot(ac, bp), d5 = dot(ab, cp), d6 = dot(ac, cp);
    const float vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
    const float den = __fdividef(1.f, va + vb + vc);
    float v = vb * den, w = vc * den;                                        // interior: a + ab v + ac w
    { const bool k = va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f; const float ww = __fdividef(d4 - d3, (d4 - d3) + (d5 - d6)); v = k ? 1.f - ww : v; w = k ? ww : w; }
    { const bool k = vb <= 0.f && d2 >= 0.f && d6 <= 0.f; const float ww = __fdividef(d2, d2 - d6); v = k ? 0.f : v; w = k ? ww : w; }
    { const bool k = d6 >= 0.f && d5 <= d6; v = k ? 0.f : v; w = k ? 1.f : w; }
    { const bool k = vc <= 0.f && d1 >= 0.f && d3 <= 0.f; const float vv = __fdividef(d1, d1 - d3); v = k ? vv : v; w = k ? 0.f : w; }
    { const bool k = d3 >= 0.f && d4 <= d3; v = k ? 1.f : v; w = k ? 0.f : w; }
    { const bool k = d1 <= 0.f && d2 <= 0.f; v = k ? 0.f : v; w = k ? 0.f : w; }
    return a + ab * v + ac * w;
}

// Exact squared distance of triangles P = (0, p1, p2) and Q = (q0, q1, q2) with the closest points (same shifted frame).
__device__ __forceinline__ float tri_tri_dist2(V3 p1, V3 p2, V3 q0, V3 q1, V3 q2, V3& cp, V3& cq)
{
    const V3 P[3] = {v3(0.f, 0.f, 0.f), p1, p2};
    const V3 Q[3] = {q0, q1, q2};
    float best = 3.0e38f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const V3 pa = P[i], pd = P[(i + 1) % 3] - P[i];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            V3 c1, c2;
            const float d = seg_seg(pa, pd, Q[j], Q[(j + 1) % 3] - Q[j], c1, c2);
            if (d < best) { best = d; cp = c1; cq = c2; }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const V3 c = closest_on_triangle(P[i], q0, q1, q2);
        const V3 d = P[i] - c;
        const float dd = dot(d, d);
        if (dd < best) { best = dd; cp = P[i]; cq = c; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const V3 c = closest_on_triangle(Q[i], P[0], p1, p2);
        const V3 d = Q[i] - c;
        const float dd = dot(d, d);
        if (dd < best) { best = dd; cp = c; cq = Q[i]; }
    }
    // interpenetrating triangles (an edge of one through the interior of the other) have no zero among the 15 candidates
    if (best > 0.f && tri_contact_normals(p1, p2, q0, q1, q2, 0.f) && tri_contact_edges(p1, p2, q0, q1, q2, 0.f)) { best = 0.f; cq = cp; }
    return best;
}

// lower bound of the squared distance between box A {ca, ea} (identity frame) and box B {cb, eb} placed by (R, t)
__device__ __forceinline__ float box_gap2(const float* __restrict__ R, const float* __restrict__ t, float4 ac, float4 ae, float4 bc, float4 be)
{
    const float slack = 2e-6f;
    const float T0 = fmaf(R[0], bc.x, fmaf(R[1], bc.y, fmaf(R[2], bc.z, t[0]))) - ac.x;
    const float T1 = fmaf(R[3], bc.x, fmaf(R[4], bc.y, fmaf(R[5], bc.z, t[1]))) - ac.y;
    const float T2 = fmaf(R[6], bc.x, fmaf(R[7], bc.y, fmaf(R[8], bc.z, t[2]))) - ac.z;
    const float f00 = fabsf(R[0]), f01 = fabsf(R[1]), f02 = fabsf(R[2]);
    const float f10 = fabsf(R[3]), f11 = fabsf(R[4]), f12 = fabsf(R[5]);
    const float f20 = fabsf(R[6]), f21 = fabsf(R[7]), f22 = fabsf(R[8]);
    const float g0 = fmaxf(0.f, fabsf(T0) - (ae.x + be.x * f00 + be.y * f01 + be.z * f02) - slack);
    const float g1 = fmaxf(0.f, fabsf(T1) - (ae.y + be.x * f10 + be.y * f11 + be.z * f12) - slack);
    const float g2 = fmaxf(0.f, fabsf(T2) - (ae.z + be.x * f20 + be.y * f21 + be.z * f22) - slack);
    const float h0 = fmaxf(0.f, fabsf(T0 * R[0] + T1 * R[3] + T2 * R[6]) - (be.x + ae.x * f00 + ae.y * f10 + ae.z * f20) - slack);
    const float h1 = fmaxf(0.f, fabsf(T0 * R[1] + T1 * R[4] + T2 * R[7]) - (be.y + ae.x * f01 + ae.y * f11 + ae.z * f21) - slack);
    const float h2 = fmaxf(0.f, fabsf(T0 * R[2] + T1 * R[5] + T2 * R[8]) - (be.z + ae.x * f02 + ae.y * f12 + ae.z * f22) - slack);
    return fmaxf(g0 * g0 + g1 * g1 + g2 * g2, h0 * h0 + h1 * h1 + h2 * h2);
}

__device__ __forceinline__ float aabb_gap2(const float4 ac, const float4 ae, const float4 bc, const float4 be)
{
    const float g0 = fmaxf(0.f, fabsf(ac.x - bc.x) - (ae.x + be.x));
    const float g1 = fmaxf(0.f, fabsf(ac.y - bc.y) - (ae.y + be.y));
    const float g2 = fmaxf(0.f, fabsf(ac.z - bc.z) - (ae.z + be.z));
    return g0 * g0 + g1 * g1 + g2 * g2;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
cc_distance_kernel(const __grid_constant__ Params P, const __grid_constant__ DistOut D)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    unsigned char* wbase = smem_raw + wib * kDistWarpSmemBytes;
    float* slot_rt = (float*)wbase;                               // [kNumSlots][24]: B into A, then A into B (unused half kept for layout)
    int32_t* slot_pair = (int32_t*)(slot_rt + kNumSlots * 24);    // [kNumSlots] reference pair index
    int32_t* slot_la = slot_pair + kNumSlots;                     // [kNumSlots] plan slot of link a (to place the closest points)
    uint2* stack = (uint2*)(slot_la + kNumSlots);                 // [kDistStackCap] {nodeA | slot << 26, nodeB}
    uint2* triq = stack + kDistStackCap;                          // [kDistTriQCap] {triA | slot << 26, triB}
    uint32_t* cbuf = (uint32_t*)(triq + kDistTriQCap);            // [32] candidate device pair indices
    const unsigned lt_mask = (1u << lane) - 1u;
    const int ndyn = P.ndyn;
    const bool approx = D.rel_err > 0.f || D.abs_err > 0.f;
    const float inv_rel = 1.f / (1.f + fmaxf(D.rel_err, 0.f));
    float best_init = D.dist_max > 0.f ? D.dist_max * D.dist_max : 3.0e38f;
    float thr_init = best_init;
    if (D.bound > 0.f && D.bound * D.bound < best_init) {
        best_init = D.bound * D.bound;
        thr_init = best_init;
        if (approx) { const float th = fmaxf(0.f, fminf(D.bound - D.abs_err, D.bound * inv_rel)); thr_init = th * th; }     // as if found here
    }

    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&P.counters[0], 1);
        c = __shfl_sync(kFull, c, 0);
        const int w = c;
        int k_begin = (P.flags & kFlagEnv) ? 0 : P.nenv_dev;
        int k_end = (P.flags & kFlagSelf) ? P.npairs_dev : P.nenv_dev;
        if (D.per_pair) {
            if (w >= P.n * P.npairs_dev) break;
            c = w / P.npairs_dev;
            k_begin = w - c * P.npairs_dev;
            k_end = k_begin + 1;
        } else if (c >= P.n) break;
        const float* xf = P.xf32 + (size_t)c * ndyn * 12;
        const float* ab = P.aabb + (size_t)c * ndyn * 8;
        constexpr int xs = 12, as = 8;

        float best2 = best_init;             // warp-uniform
        float thr2 = thr_init;               // squared pruning threshold: best2 itself, or tightened by the tolerances
        int best_pair = -1, best_la = 0;
        V3 best_p = v3(0.f, 0.f, 0.f), best_q = best_p;     // closest points in link a's frame (warp-uniform copies)

        auto pair_bound = [&](int k) -> float {
            const uint32_t pr = __ldg(&P.pairs[k]);
            const int sa = pr & 0xffff, sb = pr >> 16;
            const float4* pa = (const float4*)(sa < ndyn ? ab + as * sa : P.st_aabb + 8 * (size_t)(sa - ndyn));
            const float4* pb = (const float4*)(sb < ndyn ? ab + as * sb : P.st_aabb + 8 * (size_t)(sb - ndyn));
            return aabb_gap2(pa[0], pa[1], pb[0], pb[1]);
        };

        // ---- the most promising pair first
        int first_k = -1;
        {
            float lb = 3.0e38f; int kk = -1;
            for (int k = k_begin + lane; k < k_end; k += 32) { const float b = pair_bound(k); if (b < lb) { lb = b; kk = k; } }
            const unsigned key = __reduce_min_sync(kFull, __float_as_uint(lb));
            const unsigned who = __ballot_sync(kFull, __float_as_uint(lb) == key && kk >= 0);
            if (who) first_k = __shfl_sync(kFull, kk, __ffs(who) - 1);
        }

        int k0 = k_begin;
        bool did_first = first_k < 0;
        while (!did_first || k0 < k_end) {
            // ---- gather up to 32 root pairs whose AABB bound is still below best2
            int ncb = 0;
            if (!did_first) { if (lane == 0) cbuf[0] = (uint32_t)first_k; ncb = 1; did_first = true; }
            else {
                while (k0 < k_end && ncb == 0) {
                    const int k = k0 + lane;
                    const bool want = k < k_end && k != first_k && pair_bound(k) < thr2;
                    const unsigned m = __ballot_sync(kFull, want);
                    if (want) cbuf[__popc(m & lt_mask)] = (uint32_t)k;
                    ncb = __popc(m);
                    k0 += 32;
                }
                if (ncb == 0) break;
            }
            __syncwarp();
            int sp = 0, ntri = 0;
            {
                bool cand = false;
                uint32_t ra = 0, rb = 0;
                if (lane < ncb) {
                    const uint32_t k = cbuf[lane];
                    const uint32_t pr = __ldg(&P.pairs[k]);
                    const int la = pr & 0xffff, lb = pr >> 16;
                    float A[12], Bx[12], R[9], t[3];
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
                    float* rt = slot_rt + 24 * lane;
#pragma unroll
                    for (int i = 0; i < 9; ++i) rt[i] = R[i];
                    rt[9] = t[0]; rt[10] = t[1]; rt[11] = t[2];
                    slot_pair[lane] = __ldg(&P.pair_orig[k]);
                    slot_la[lane] = la;
                    ra = (uint32_t)__ldg(&P.slot_root[la]); rb = (uint32_t)__ldg(&P.slot_root[lb]);
                    cand = true;
                }
                // first items: two leaves go straight to the triangle queue, otherwise the larger root box is split
                // (item = the two children of the split node x the other node; bit 31 of y: the children live in tree B)
                bool leafpair = false;
                uint2 it0 = make_uint2(0u, 0u);
                if (cand) {
                    const float4 a0 = __ldg(&P.nodes[2 * ra]), a1 = __ldg(&P.nodes[2 * ra + 1]);
                    const float4 b0 = __ldg(&P.nodes[2 * rb]), b1 = __ldg(&P.nodes[2 * rb + 1]);
                    const int la = __float_as_int(a0.w), lb = __float_as_int(b0.w);
                    const uint32_t sb = (uint32_t)lane << 26;
                    if (la < 0 && lb < 0) { leafpair = true; it0 = make_uint2((uint32_t)(~la) | sb, (uint32_t)(~lb)); }
                    else if (lb < 0 || (la >= 0 && a1.w > b1.w)) it0 = make_uint2((uint32_t)la | sb, rb);
                    else it0 = make_uint2((uint32_t)lb | sb, ra | 0x80000000u);
                }
                const unsigned tm0 = __ballot_sync(kFull, leafpair);
                const unsigned pm0 = __ballot_sync(kFull, cand && !leafpair);
                if (leafpair) triq[__popc(tm0 & lt_mask)] = it0;
                if (cand && !leafpair) stack[__popc(pm0 & lt_mask)] = it0;
                ntri = __popc(tm0);
                sp = __popc(pm0);
            }
            __syncwarp();

            // ---- drain: 32 items at a time (two box bounds each); leaf pairs are queued and evaluated 32 at a time
            while (sp > 0 || ntri > 0) {
                if (ntri >= 32 || (sp == 0 && ntri > 0)) {
                    const int tn = ntri < 32 ? ntri : 32;
                    float d2 = 3.0e38f;
                    V3 cp = v3(0.f, 0.f, 0.f), cq = cp;
                    uint32_t s = 0;
                    if (lane < tn) {
                        const uint2 e = triq[ntri - 1 - lane];
                        s = e.x >> 26;
                        const uint32_t ta_idx = e.x & 0x3ffffffu;
                        const float4 v0 = __ldg(P.tris + 3 * (size_t)ta_idx);
                        V3 p1, p2, q0, q1, q2;
                        load_tri_pair(P.tris, slot_rt + 24 * s, ta_idx, e.y, p1, p2, q0, q1, q2);
                        d2 = tri_tri_dist2(p1, p2, q0, q1, q2, cp, cq);
                        const V3 org = v3(v0.x, v0.y, v0.z);
                        cp = cp + org; cq = cq + org;
                    }
                    ntri -= tn;
                    const unsigned key = __reduce_min_sync(kFull, __float_as_uint(d2));
                    if (__uint_as_float(key) < best2) {
                        const unsigned who = __ballot_sync(kFull, __float_as_uint(d2) == key);
                        const int src = __ffs(who) - 1;
                        best2 = __uint_as_float(key);
                        thr2 = best2;
                        if (approx && best2 < 1.0e37f) {
                            const float b = sqrtf(best2);
                            const float th = fmaxf(0.f, fminf(b - D.abs_err, b * inv_rel));
                            thr2 = th * th;
                        }
                        const uint32_t bs = __shfl_sync(kFull, s, src);
                        best_pair = slot_pair[bs];
                        best_la = slot_la[bs];
                        best_p = v3(__shfl_sync(kFull, cp.x, src), __shfl_sync(kFull, cp.y, src), __shfl_sync(kFull, cp.z, src));
                        best_q = v3(__shfl_sync(kFull, cq.x, src), __shfl_sync(kFull, cq.y, src), __shfl_sync(kFull, cq.z, src));
                    }
                    __syncwarp();
                    if (best2 == 0.f) { sp = 0; ntri = 0; }       // touching / penetrating: nothing can be smaller
                    continue;
                }
                // near the capacity only one item is expanded per iteration (a single path cannot overflow the margin)
                const int width = sp > kDistStackCap - 160 ? 1 : 32;
                const int n = sp < width ? sp : width;
                const bool valid = lane < n;
                uint2 it = make_uint2(0u, 0u);
                if (valid) it = stack[sp - 1 - lane];
                sp -= n;
                __syncwarp();
                // each item: children x, x+1 of one tree against node y of the other; a child whose bound reaches the running
                // minimum is dropped here and never stored
                bool lf0 = false, lf1 = false, ps0 = false, ps1 = false;
                uint2 o0 = make_uint2(0u, 0u), o1 = o0;
                float g0 = 3.0e38f, g1 = 3.0e38f;
                if (valid) {
                    const uint32_t sbits = it.x & 0xfc000000u;
                    const uint32_t x = it.x & 0x3ffffffu, y = it.y & 0x7fffffffu;
                    const bool side = (it.y >> 31) != 0u;               // false: the children are in tree A, true: in tree B
                    const float4 c0a = __ldg(&P.nodes[2 * x]), c0b = __ldg(&P.nodes[2 * x + 1]);
                    const float4 c1a = __ldg(&P.nodes[2 * x + 2]), c1b = __ldg(&P.nodes[2 * x + 3]);
                    const float4 oa = __ldg(&P.nodes[2 * y]), ob = __ldg(&P.nodes[2 * y + 1]);
                    const float* rt = slot_rt + 24 * (it.x >> 26);
                    float R[9], t[3];
#pragma unroll
                    for (int k = 0; k < 9; ++k) R[k] = rt[k];
                    t[0] = rt[9]; t[1] = rt[10]; t[2] = rt[11];
                    g0 = side ? box_gap2(R, t, oa, ob, c0a, c0b) : box_gap2(R, t, c0a, c0b, oa, ob);
                    g1 = side ? box_gap2(R, t, oa, ob, c1a, c1b) : box_gap2(R, t, c1a, c1b, oa, ob);
                    const int lo = __float_as_int(oa.w);
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        if ((ch ? g1 : g0) < thr2) {
                            const int lc = __float_as_int(ch ? c1a.w : c0a.w);
                            const float sc = ch ? c1b.w : c0b.w;
                            uint2 out;
                            const bool tq = lc < 0 && lo < 0;
                            if (tq) out = side ? make_uint2((uint32_t)(~lo) | sbits, (uint32_t)(~lc)) : make_uint2((uint32_t)(~lc) | sbits, (uint32_t)(~lo));
                            else if (lo < 0 || (lc >= 0 && sc > ob.w)) out = make_uint2((uint32_t)lc | sbits, it.y);                          // split the child
                            else out = make_uint2((uint32_t)lo | sbits, (x + ch) | (side ? 0u : 0x80000000u));                                // split the other
                            if (ch) { o1 = out; lf1 = tq; ps1 = !tq; } else { o0 = out; lf0 = tq; ps0 = !tq; }
                        }
                    }
                    // the nearer child is stored last, so that it is expanded first
                    if (ps0 && ps1 && g0 < g1) { const uint2 tmp = o0; o0 = o1; o1 = tmp; }
                }
                {
                    const unsigned l0 = __ballot_sync(kFull, lf0), l1 = __ballot_sync(kFull, lf1);
                    if (lf0) triq[ntri + __popc(l0 & lt_mask)] = o0;
                    if (lf1) triq[ntri + __popc(l0) + __popc(l1 & lt_mask)] = o1;
                    ntri += __popc(l0) + __popc(l1);
                    const unsigned p0 = __ballot_sync(kFull, ps0), p1 = __ballot_sync(kFull, ps1);
                    // (writing the outputs back in reverse lane order — the change worth +18 .. 37 % in the collision kernel — is
                    // neutral here: 1.477 vs 1.473 M exact queries/s, profiles/r02g_ab_dist_push.log; the nearer-child-last rule
                    // below already keeps the most promising item of each lane on top)
                    if (ps0) stack[sp + __popc(p0 & lt_mask)] = o0;
                    if (ps1) stack[sp + __popc(p0) + __popc(p1 & lt_mask)] = o1;
                    sp += __popc(p0) + __popc(p1);
                }
                __syncwarp();
            }
            if (best2 == 0.f) break;
        }

        if (lane == 0) {
            double dist = best_pair >= 0 ? (double)sqrtf(best2) : (D.dist_max > 0.f ? (double)D.dist_max : 1e30);
            D.min_distance[w] = dist;
            if (D.min_pair) D.min_pair[w] = best_pair;
            if (D.points) {
                double* o = D.points + 6 * (size_t)w;
                if (best_pair >= 0) {
                    const float* A = best_la < ndyn ? xf + xs * best_la : P.st_xf + 12 * (size_t)(best_la - ndyn);
                    const V3 pts[2] = {best_p, best_q};
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        o[3 * k + 0] = (double)(A[0] * pts[k].x + A[1] * pts[k].y + A[2] * pts[k].z + A[9]) + P.center[0];
                        o[3 * k + 1] = (double)(A[3] * pts[k].x + A[4] * pts[k].y + A[5] * pts[k].z + A[10]) + P.center[1];
                        o[3 * k + 2] = (double)(A[6] * pts[k].x + A[7] * pts[k].y + A[8] * pts[k].z + A[11]) + P.center[2];
                    }
                } else {
                    for (int k = 0; k < 6; ++k) o[k] = 0.0;
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace ccdev
