// cudacollision.cu — host side of libcudacollision.so: the C ABI of include/cudacollision.h.
//
// Host responsibilities (all one-off or per-batch, never per-configuration):
//   * keep the flat scene the OpenRAVE-side plugin pushes in (bodies, link meshes, poses, kinematics);
//   * build one link-frame AABB BVH per link (cc_scene_commit) and keep nodes/triangles resident in HBM;
//   * bind a query: lay out the pair table in pqprave's iteration order (collisionPQP.h:432-498, 364-386),
//     hoist the pairs whose relative pose cannot change with the active DOFs, compile the FK chain program;
//   * launch ccdev::cc_check_kernel (cc_device.cuh) on the caller's stream.
// There is NO CPU collision path here: without a CUDA device every query entry point fails with CC_ERR_CUDA.
#include <cuda_runtime.h>
#include <cstdlib>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <limits>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <functional>
#include <vector>

#include "../../include/cudacollision.h"
#include "cc_device.cuh"
#include "cc_distance.cuh"
#include "cc_analytic.cuh"
#include "cc_mimic_expr.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CC_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(CC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
    } while (0)

// ------------------------------------------------------------------------------------------------ host scene
// matches the device's 2 x float4; the right child is always left + 1, so the last word carries |e|^2 (the "size" the
// descent rule compares) instead
struct HNode { float c[3]; int32_t left; float e[3]; float size2; };
static_assert(sizeof(HNode) == 32, "node layout");

struct HLink {
    std::string name;
    std::vector<double> verts;     // xyz
    std::vector<int32_t> tris;
    bool enabled = true;
    double xf[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    bool dirty = false;            // mesh changed since the last commit
    std::vector<HNode> nodes;      // local indices
    std::vector<float> tridata;    // 12 floats per triangle
    int node_base = -1, tri_base = -1;
    std::vector<cc_geom> prims;    // analytic primitives of the link (cc_body_set_link_primitives); used in analytic mode only
    int ntris() const { return (int)tris.size() / 3; }
};

struct HBody {
    std::string name;
    bool is_robot = false;
    bool removed = false;
    std::vector<HLink> links;
    std::vector<int> attached;
    int ndof = 0;
    std::vector<cc_joint> joints;
    std::vector<uint8_t> affects;      // [ndof][nlinks]
    std::vector<double> dofvals;
    std::vector<int32_t> nonadjacent;
    // KinBody::_vGrabbedBodies in grab order (kinbodygrab.cpp): the grabbed body's links ride on `link` of this body
    struct Grab {
        int body = -1, link = -1;
        std::vector<double> rel;            // [links of the grabbed body][8], pose of each relative to the grabbing link
        std::vector<int32_t> noncolliding;  // (body, link) pairs: Grabbed::_listNonCollidingLinksWhenGrabbed
    };
    std::vector<Grab> grabs;
    // equations of CC_MIMIC_EXTERNAL joints handed over with cc_body_set_mimic_equation: [njoints], empty where there is none
    std::vector<ccexpr::Program> mimic_prog;
};

const float kBoxInflate = 1e-5f;   // metres added to every half extent (conservative pruning, see cc_device.cuh)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t reserve(size_t n)
    {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    cudaError_t upload(const std::vector<T>& v, cudaStream_t st = 0)
    {
        cudaError_t e = reserve(v.size() ? v.size() : 1);
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    }
};

}  // namespace

constexpr int kMaxLanes = 4;      // pipeline lanes of a big batch (streams + scratch sets); CC_TUNE_LANES = 2 .. 4 picks how many are used

struct cc_scene {
    int device = 0;
    int num_sms = 0;
    std::vector<HBody*> bodies;
    bool committed = false;
    DevBuf<float4> d_nodes, d_tris;
    int total_nodes = 0, total_tris = 0;
    DevBuf<uint2> d_spill[kMaxLanes];       // one per pipeline lane (one chunk per lane is in flight at a time)
    int spill_warps = 0;
    cudaStream_t lane_stream[kMaxLanes] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxLanes] = {};
    int32_t* h_err = nullptr;               // error flag (a traversal overflowed its spill area): page-locked, mapped into the
    int32_t* d_err = nullptr;               //   device address space, so the host reads it after a sync without a copy
    int trav_blocks_per_sm = 0;             // occupancy of cc_traverse_kernel (queried once)
    DevBuf<unsigned long long> d_stats;     // [0] bv, [1] tri
    bool counters = false;
    cc_stats stats = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_pending = false;
    float tri_eps = 0.0f;
    bool analytic = false;                  // fclrave semantics: links with primitives collide as exact convex shapes (cc_analytic.cuh)
    bool pair_tables = true;                // certified free-space tables for pairs that depend on <= 2 active DOFs
    // multi-device scene (cc_scene_create_multi): this is the replica on devices[0]; `peers` are full replicas on the other
    // devices.  Every mutating entry point is mirrored onto them in the same order, so body / link / pair indices agree.
    std::vector<cc_scene*> peers;
    // per-chunk timeline of the last batched call (cc_scene_trace_timeline): 3 timed events per chunk on the chunk's lane stream
    bool trace_timeline = false;
    std::vector<cudaEvent_t> tl_events;
    std::vector<int> tl_lane;               // lane of every chunk of the last traced call
    uint64_t pose_epoch = 0;                // bumped by cc_body_set_link_transforms / cc_body_set_dof_values: a bound query whose
                                            // snapshot is older than this must not prove anything from the live scene state
    ~cc_scene()
    {
        for (cc_scene* p : peers) { cudaSetDevice(p->device); cudaDeviceSynchronize(); delete p; }
        cudaSetDevice(device);
        for (HBody* b : bodies) delete b;
        for (cudaEvent_t e : tl_events) cudaEventDestroy(e);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (h_err) cudaFreeHost(h_err);
        for (int i = 0; i < kMaxLanes; ++i) { if (ev_join[i]) cudaEventDestroy(ev_join[i]); if (lane_stream[i]) cudaStreamDestroy(lane_stream[i]); }
    }
};

namespace {

// ------------------------------------------------------------------------------------------------ BVH build
struct Builder {
    const HLink& L;
    std::vector<HNode>& nodes;
    std::vector<int> order;                 // triangle permutation
    std::vector<double> cen;                // centroids
    Builder(const HLink& l, std::vector<HNode>& n) : L(l), nodes(n) {}

    void bounds(int first, int count, double mn[3], double mx[3]) const
    {
        for (int k = 0; k < 3; ++k) { mn[k] = 1e300; mx[k] = -1e300; }
        for (int i = first; i < first + count; ++i) {
            int t = order[i];
            for (int v = 0; v < 3; ++v) {
                const double* p = &L.verts[3 * (size_t)L.tris[3 * t + v]];
                for (int k = 0; k < 3; ++k) { if (p[k] < mn[k]) mn[k] = p[k]; if (p[k] > mx[k]) mx[k] = p[k]; }
            }
        }
    }

    void fill_box(HNode& n, const double mn[3], const double mx[3]) const
    {
        for (int k = 0; k < 3; ++k) {
            double c = 0.5 * (mn[k] + mx[k]);
            float cf = (float)c;
            double e = 0.5 * (mx[k] - mn[k]) + fabs(c - (double)cf) + (double)kBoxInflate;
            float ef = (float)e;
            if ((double)ef < e) ef = nextafterf(ef, std::numeric_limits<float>::infinity());
            n.c[k] = cf; n.e[k] = ef;
        }
        // what "the larger box" means when the descent picks the node to split: the volume (default; +1.3 .. 2.3 % over |e|^2 on the
        // three robot workloads, profiles/r01C_ab_bvh_knobs.log); CC_TUNE_BVH_SIZE = 0 |e|^2, 1 area, 3 longest extent
        static const int size_kind = getenv("CC_TUNE_BVH_SIZE") ? atoi(getenv("CC_TUNE_BVH_SIZE")) : 2;
        n.size2 = size_kind == 1 ? n.e[0] * n.e[1] + n.e[1] * n.e[2] + n.e[2] * n.e[0]
                : size_kind == 2 ? n.e[0] * n.e[1] * n.e[2]
                : size_kind == 3 ? std::max(n.e[0], std::max(n.e[1], n.e[2]))
                : n.e[0] * n.e[0] + n.e[1] * n.e[1] + n.e[2] * n.e[2];
    }

    void build(int node, int first, int count)
    {
        double mn[3], mx[3];
        bounds(first, count, mn, mx);
        fill_box(nodes[node], mn, mx);
        if (count == 1) {
            nodes[node].left = ~order[first];
            return;
        }
        // split: binned surface-area heuristic over the three axes (the verdict does not depend on the tree's shape, only the
        // number of box tests does); CC_TUNE_BVH=mid keeps the plain midpoint of the longest centroid axis; the median is the
        // fallback of both
        double cmn[3] = {1e300, 1e300, 1e300}, cmx[3] = {-1e300, -1e300, -1e300};
        for (int i = first; i < first + count; ++i)
            for (int k = 0; k < 3; ++k) { double c = cen[3 * (size_t)order[i] + k]; if (c < cmn[k]) cmn[k] = c; if (c > cmx[k]) cmx[k] = c; }
        int ax = 0;
        if (cmx[1] - cmn[1] > cmx[ax] - cmn[ax]) ax = 1;
        if (cmx[2] - cmn[2] > cmx[ax] - cmn[ax]) ax = 2;
        double mid = 0.5 * (cmn[ax] + cmx[ax]);
        static const bool use_sah = !(getenv("CC_TUNE_BVH") && !strcmp(getenv("CC_TUNE_BVH"), "mid"));
        if (use_sah && count > 2) {
            constexpr int NBMAX = 64;
            static const int NB = getenv("CC_TUNE_BVH_BINS") ? std::min(NBMAX, std::max(2, atoi(getenv("CC_TUNE_BVH_BINS")))) : 16;
            static const int cost_kind = getenv("CC_TUNE_BVH_COST") ? atoi(getenv("CC_TUNE_BVH_COST")) : 0;    // 0 area, 1 volume, 2 |e|^2, 3 sum of extents
            auto measure = [](double dx, double dy, double dz) {
                switch (cost_kind) {
                    case 1: return dx * dy * dz;
                    case 2: return dx * dx + dy * dy + dz * dz;
                    case 3: return dx + dy + dz;
                    default: return dx * dy + dy * dz + dz * dx;
                }
            };
            double best_cost = 1e300;
            for (int a = 0; a < 3; ++a) {
                const double w = cmx[a] - cmn[a];
                if (!(w > 0)) continue;
                double bmn[NBMAX][3], bmx[NBMAX][3];
                int bcnt[NBMAX];
                for (int j = 0; j < NB; ++j) { bcnt[j] = 0; for (int k = 0; k < 3; ++k) { bmn[j][k] = 1e300; bmx[j][k] = -1e300; } }
                const double sc = NB / w;
                for (int i = first; i < first + count; ++i) {
                    const int t = order[i];
                    int j = (int)((cen[3 * (size_t)t + a] - cmn[a]) * sc);
                    if (j >= NB) j = NB - 1;
                    bcnt[j]++;
                    for (int v = 0; v < 3; ++v) {
                        const double* p = &L.verts[3 * (size_t)L.tris[3 * t + v]];
                        for (int k = 0; k < 3; ++k) { if (p[k] < bmn[j][k]) bmn[j][k] = p[k]; if (p[k] > bmx[j][k]) bmx[j][k] = p[k]; }
                    }
                }
                // suffix areas, then a prefix sweep
                double rarea[NBMAX];
                int rcnt[NBMAX];
                {
                    double amn[3] = {1e300, 1e300, 1e300}, amx[3] = {-1e300, -1e300, -1e300};
                    int cacc = 0;
                    for (int j = NB - 1; j >= 1; --j) {
                        cacc += bcnt[j];
                        for (int k = 0; k < 3; ++k) { amn[k] = std::min(amn[k], bmn[j][k]); amx[k] = std::max(amx[k], bmx[j][k]); }
                        const double dx = amx[0] - amn[0], dy = amx[1] - amn[1], dz = amx[2] - amn[2];
                        rarea[j] = cacc ? measure(dx, dy, dz) : 0.0;
                        rcnt[j] = cacc;
                    }
                }
                double amn[3] = {1e300, 1e300, 1e300}, amx[3] = {-1e300, -1e300, -1e300};
                int cacc = 0;
                for (int j = 0; j + 1 < NB; ++j) {
                    cacc += bcnt[j];
                    for (int k = 0; k < 3; ++k) { amn[k] = std::min(amn[k], bmn[j][k]); amx[k] = std::max(amx[k], bmx[j][k]); }
                    if (cacc == 0 || rcnt[j + 1] == 0) continue;
                    const double dx = amx[0] - amn[0], dy = amx[1] - amn[1], dz = amx[2] - amn[2];
                    const double cost = measure(dx, dy, dz) * cacc + rarea[j + 1] * rcnt[j + 1];
                    if (cost < best_cost) { best_cost = cost; ax = a; mid = cmn[a] + (j + 1) * (w / NB); }
                }
            }
        }
        int* b = &order[first];
        int nleft = (int)(std::partition(b, b + count, [&](int t) { return cen[3 * (size_t)t + ax] < mid; }) - b);
        if (nleft == 0 || nleft == count) {
            nleft = count / 2;
            std::nth_element(b, b + nleft, b + count, [&](int x, int y) {
                double cx = cen[3 * (size_t)x + ax], cy = cen[3 * (size_t)y + ax];
                return cx < cy || (cx == cy && x < y);
            });
        }
        int l = (int)nodes.size();
        nodes.push_back(HNode()); nodes.push_back(HNode());
        nodes[node].left = l;
        // the descent pops child 0 first: CC_TUNE_BVH_ORDER picks which half that is (0 = as partitioned, 1 = larger volume,
        // 2 = smaller volume, 3 = more triangles)
        static const int order_kind = getenv("CC_TUNE_BVH_ORDER") ? atoi(getenv("CC_TUNE_BVH_ORDER")) : 0;
        bool swap_halves = false;
        if (order_kind == 3) swap_halves = count - nleft > nleft;
        else if (order_kind == 1 || order_kind == 2) {
            double amn[3], amx[3], bmn[3], bmx[3];
            bounds(first, nleft, amn, amx);
            bounds(first + nleft, count - nleft, bmn, bmx);
            const double va = (amx[0] - amn[0] + 2e-5) * (amx[1] - amn[1] + 2e-5) * (amx[2] - amn[2] + 2e-5);
            const double vb = (bmx[0] - bmn[0] + 2e-5) * (bmx[1] - bmn[1] + 2e-5) * (bmx[2] - bmn[2] + 2e-5);
            swap_halves = order_kind == 1 ? vb > va : vb < va;
        }
        if (swap_halves) {
            build(l, first + nleft, count - nleft);
            build(l + 1, first, nleft);
        } else {
            build(l, first, nleft);
            build(l + 1, first + nleft, count - nleft);
        }
    }
};

void build_link_bvh(HLink& L)
{
    int nt = L.ntris();
    L.nodes.clear(); L.tridata.clear();
    if (nt == 0) return;
    L.nodes.reserve(2 * (size_t)nt);
    L.nodes.push_back(HNode());
    Builder B(L, L.nodes);
    B.order.resize(nt);
    B.cen.resize(3 * (size_t)nt);
    for (int t = 0; t < nt; ++t) {
        B.order[t] = t;
        for (int k = 0; k < 3; ++k)
            B.cen[3 * (size_t)t + k] = (L.verts[3 * (size_t)L.tris[3 * t] + k] + L.verts[3 * (size_t)L.tris[3 * t + 1] + k] +
                                        L.verts[3 * (size_t)L.tris[3 * t + 2] + k]) / 3.0;
    }
    B.build(0, 0, nt);
    L.tridata.resize(12 * (size_t)nt);
    for (int t = 0; t < nt; ++t)
        for (int v = 0; v < 3; ++v) {
            const double* p = &L.verts[3 * (size_t)L.tris[3 * t + v]];
            float* o = &L.tridata[12 * (size_t)t + 4 * v];
            o[0] = (float)p[0]; o[1] = (float)p[1]; o[2] = (float)p[2]; o[3] = 0.0f;
        }
}

// quaternion pose -> rotation matrix (row major), double
void quat_to_matrix(const double* x, double* Rm)
{
    const double q0 = x[0], q1 = x[1], q2 = x[2], q3 = x[3];
    const double il = 2.0 / (q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    Rm[0] = 1 - il * (q2 * q2 + q3 * q3); Rm[1] = il * (q1 * q2 - q0 * q3); Rm[2] = il * (q1 * q3 + q0 * q2);
    Rm[3] = il * (q1 * q2 + q0 * q3); Rm[4] = 1 - il * (q1 * q1 + q3 * q3); Rm[5] = il * (q2 * q3 - q0 * q1);
    Rm[6] = il * (q1 * q3 - q0 * q2); Rm[7] = il * (q2 * q3 + q0 * q1); Rm[8] = 1 - il * (q1 * q1 + q2 * q2);
}

// half extents of a primitive's bounding box in its own frame (box: itself; cylinder r, r, h/2; capsule r, r, h/2 + r)
void prim_half_extents(const cc_geom& g, double e[3])
{
    if (g.type == CC_GEOM_BOX) { e[0] = g.dims[0]; e[1] = g.dims[1]; e[2] = g.dims[2]; }
    else if (g.type == CC_GEOM_SPHERE) { e[0] = e[1] = e[2] = g.dims[0]; }
    else if (g.type == CC_GEOM_CYLINDER) { e[0] = e[1] = g.dims[0]; e[2] = 0.5 * g.dims[1]; }
    else { e[0] = e[1] = g.dims[0]; e[2] = 0.5 * g.dims[1] + g.dims[0]; }
}

// The root box of a link bounds its tessellated mesh; the exact primitives stick out of their tessellation by the sagitta, so
// the root box (link boxes, static world boxes: all pruning starts there) is grown to hold them too.
void expand_root_for_primitives(HLink& L)
{
    if (L.prims.empty() || L.nodes.empty()) return;
    double mn[3], mx[3];
    for (int k = 0; k < 3; ++k) { mn[k] = (double)L.nodes[0].c[k] - (double)L.nodes[0].e[k]; mx[k] = (double)L.nodes[0].c[k] + (double)L.nodes[0].e[k]; }
    for (const cc_geom& g : L.prims) {
        double Rm[9], e[3];
        quat_to_matrix(g.pose, Rm);
        prim_half_extents(g, e);
        for (int r = 0; r < 3; ++r) {
            double ext = 0;
            for (int k = 0; k < 3; ++k) ext += fabs(Rm[3 * r + k]) * e[k];
            ext += 1e-5;
            mn[r] = std::min(mn[r], g.pose[4 + r] - ext); mx[r] = std::max(mx[r], g.pose[4 + r] + ext);
        }
    }
    for (int k = 0; k < 3; ++k) {
        const double c = 0.5 * (mn[k] + mx[k]);
        const float cf = (float)c;
        const double e = 0.5 * (mx[k] - mn[k]) + fabs(c - (double)cf);
        float ef = (float)e;
        if ((double)ef < e) ef = nextafterf(ef, std::numeric_limits<float>::infinity());
        L.nodes[0].c[k] = cf; L.nodes[0].e[k] = ef;
    }
}

// quaternion pose -> FP32 R|t relative to `center` (matrixFromQuat, geometry.h:779-800)
void pose_to_rt32(const double* x, const double center[3], float* o)
{
    double q0 = x[0], q1 = x[1], q2 = x[2], q3 = x[3];
    double il = 2.0 / (q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    double qq1 = il * q1 * q1, qq2 = il * q2 * q2, qq3 = il * q3 * q3;
    o[0] = (float)(1 - qq2 - qq3);            o[1] = (float)(il * (q1 * q2 - q0 * q3)); o[2] = (float)(il * (q1 * q3 + q0 * q2));
    o[3] = (float)(il * (q1 * q2 + q0 * q3)); o[4] = (float)(1 - qq1 - qq3);            o[5] = (float)(il * (q2 * q3 - q0 * q1));
    o[6] = (float)(il * (q1 * q3 - q0 * q2)); o[7] = (float)(il * (q2 * q3 + q0 * q1)); o[8] = (float)(1 - qq1 - qq2);
    o[9] = (float)(x[4] - center[0]); o[10] = (float)(x[5] - center[1]); o[11] = (float)(x[6] - center[2]);
}

// A plan = one pair list laid out for the device: its slot space (dynamic query links first, then everything static),
// the FK program pruned to the links it needs, and the static poses / boxes.
struct Plan {
    std::vector<uint32_t> pairs;     // slot_a | slot_b << 16
    std::vector<int32_t> orig;
    int nenv = 0, ndyn = 0, nstatic = 0, nsave = 0, npose = 1;
    std::vector<std::pair<int, int>> statics;   // (body, link) of every static slot (poses refreshed by cc_query_refresh)
    std::vector<double> st_pose;                // [nstatic][8] the poses those slots were laid out with (what moved since? cc_query_refresh)
    std::vector<int32_t> emit;       // [nq] link -> dynamic slot or -1
    std::vector<ccdev::FkStep> steps;
    DevBuf<uint32_t> d_pairs;
    DevBuf<int32_t> d_orig, d_emit, d_slot_root;
    DevBuf<float> d_st_xf, d_st_aabb, d_link_box;
    DevBuf<ccdev::FkStep> d_steps;
    // analytic link pairs of the plan (same slot space; never part of `pairs`): evaluated by cc_analytic_kernel after the traversal
    std::vector<uint32_t> apairs;
    std::vector<int32_t> aorig;
    int anenv = 0;
    DevBuf<uint32_t> d_apairs;
    DevBuf<int32_t> d_aorig, d_slot_prim;
    DevBuf<ccdev::Prim> d_prims;
    // split traversal tables for S = 2, 4, 8 (index log2(S) - 1), built with the plan
    struct Split { DevBuf<uint32_t> d_pairs; DevBuf<int32_t> d_orig, d_part_begin, d_part_nenv; std::vector<int32_t> orig; };
    Split split[3];
};

bool body_attached(const HBody& a, int other)
{
    return std::find(a.attached.begin(), a.attached.end(), other) != a.attached.end();
}

}  // namespace

struct cc_query {
    cc_scene* scene = nullptr;
    int body = -1;
    int nq = 0, ncols = 0, nslots = 0;     // nq: query links = the body's links, then the links of every grabbed body
    std::vector<int> ql_body, ql_link;       // [nq] which (body, link) a query link is
    std::vector<int> ql_host;                // [nq] the body link whose motion it follows (itself for the body's own links)
    std::vector<int> ql_grab;                // [nq] index into HBody::grabs, -1 for the body's own links
    const HLink& qlink(int l) const;
    std::vector<int32_t> quads;              // [npairs][4] {query link, other body, other link, class}
    std::vector<int32_t> pair_slot;          // [npairs] slot of the other link
    std::vector<uint8_t> pair_live;          // [npairs] 0: can never collide (disabled / no mesh), kept for indexing only
    int nenv_total = 0, nself_total = 0;
    Plan full, mainl, hoisted;               // pose mode | batched FK mode | pose-invariant pairs (evaluated once at bind)
    // certified free-space tables of the batched FK-mode plan (built once enough configurations have gone through)
    bool allow_tables = false, tables_done = false;
    uint4* contact_out = nullptr;            // set around a cc_contacts_poses launch
    uint32_t* contact_count = nullptr;
    int contact_cap = 0;
    double cert_build_ms = 0.0;              // wall time of the last certificate build
    float cert_tri_eps = 0.0f;               // cc_scene_set_tri_epsilon value the certificates were built with
    std::vector<int> col_of_dof;             // [ndof] column of q feeding a DOF, -1: not active
    uint64_t fk_configs_seen = 0;
    uint64_t pose_epoch = 0;                 // cc_scene::pose_epoch when the query captured its snapshot (create / refresh)
    std::vector<cc_certificate> certs;       // what the certificate builds proved, pair by pair (grid certificates only)
    std::vector<uint8_t> cert_dropped;       // [npairs] certified collision-free over the whole active range: out of the batched plan
    std::vector<int32_t> table_of_orig;      // [npairs] bit table of a partly certified pair, -1: none
    std::vector<ccdev::PairTable> h_tables;  // host copies of the tables / their bits (new certificates are appended)
    std::vector<uint32_t> h_bits;
    std::vector<double> snap_static;         // poses (8 doubles per static slot of `mainl`) the certificates were proven on
    int ntables = 0;
    DevBuf<int32_t> d_pair_table;            // mainl order
    DevBuf<int32_t> d_pair_table_split[3];   // split orders
    DevBuf<ccdev::PairTable> d_tables;
    DevBuf<uint32_t> d_table_bits;
    int ndropped = 0;                        // pairs certified collision-free over the whole active range and removed
    std::vector<int> p_main;                 // original indices of the pairs of the batched FK-mode plan
    std::vector<int> p_main_all;             // the same before any certificate removed pairs
    std::vector<int> p_hoist;                // pose-invariant pairs
    std::vector<uint8_t> moves;              // [nq] link is moved by some active DOF
    Plan dist;                               // tolerance queries in FK mode: every live pair, nothing hoisted (built on first use)
    // minimum-distance queries in FK mode: the pairs between links no active DOF moves are measured ONCE (dist_static, at the
    // current pose) and every configuration's search starts from that bound; dist_main holds the rest
    Plan dist_main, dist_static;
    std::vector<int> p_dist_main, p_dist_static;
    std::vector<uint8_t> dist_dynamic;
    bool dist_main_ready = false;
    struct StaticMin { bool valid = false; double d = 0; int pair = -1; double pts[6] = {0, 0, 0, 0, 0, 0}; } dist_static_min[4];
    bool dist_ready = false;
    std::vector<int> p_full;                 // indices of the live pairs
    std::vector<int> p_analytic;             // pairs with an analytic link (scene in analytic mode): out of every traversal plan
    DevBuf<double> d_dist, d_dpts;
    DevBuf<int32_t> d_dpair;
    bool has_fk = false;
    std::vector<ccdev::FkStep> steps;       // every joint, depth-first order
    // per-chunk scratch (grown on demand, shared by the plans)
    DevBuf<float> d_xf32[kMaxLanes], d_aabb[kMaxLanes];      // one set per pipeline lane: consecutive chunks overlap on the lanes' streams
    DevBuf<uint32_t> d_cand[kMaxLanes];              // broadphase candidate bits of a chunk (cc_broad_kernel -> traversal)
    DevBuf<double> d_fksave[kMaxLanes];
    DevBuf<int32_t> d_counters[kMaxLanes];
    DevBuf<uint32_t> d_hoisted_mask;
    std::vector<uint32_t> hoisted_mask;
    int hoisted_first_env = 0x7fffffff, hoisted_first_self = 0x7fffffff;
    int mask_words = 1;
    double base[8];
    double center[3];
    std::vector<double> cur_active;          // current values of the active DOFs
    DevBuf<uint32_t> d_work;                 // optional per-configuration work trace (cc_query_get_work_trace)
    bool trace_work = false;
    int work_n = 0;
    std::vector<cc_query*> peers;            // the same query bound on every peer replica of a multi-device scene
    int affine_mask = 0, naffine = 0;        // affine active DOFs: their columns are the last ones of q
    int nexternal = 0;                       // externally evaluated mimic joints: their columns follow the active joint DOFs
    double affine_axis[4] = {0, 0, 1, 0}, affine_qstart[4] = {1, 0, 0, 0};
    // staging for the host-buffer API
    DevBuf<double> d_q, d_fk;
    DevBuf<uint32_t> d_verdict, d_mask;
    DevBuf<int32_t> d_fe, d_fs;
};

const HLink& cc_query::qlink(int l) const { return scene->bodies[ql_body[l]]->links[ql_link[l]]; }

namespace {

// Bytes of per-chunk scratch one configuration needs.
size_t scratch_bytes_per_config(const Plan& pl)
{
    return (size_t)pl.ndyn * (48 + 32) + (size_t)pl.nsave * 56 + (pl.pairs.size() + 31) / 32 * 4;
}

int ensure_scratch(cc_query* Q, const Plan& pl, int cap, int nsets)
{
    for (int k = 0; k < nsets; ++k) {
        CC_CUDA(Q->d_xf32[k].reserve((size_t)cap * std::max(1, pl.ndyn) * 12));
        CC_CUDA(Q->d_aabb[k].reserve((size_t)cap * std::max(1, pl.ndyn) * 8));
        CC_CUDA(Q->d_fksave[k].reserve((size_t)cap * std::max(1, pl.nsave) * 7));
        CC_CUDA(Q->d_counters[k].reserve(2));
        CC_CUDA(Q->d_cand[k].reserve((size_t)cap * std::max<size_t>(1, (pl.pairs.size() + 31) / 32)));
    }
    return CC_OK;
}

// Host buffers of the host-buffer entry points: when given, every chunk copies its own slice in before its FK kernel and
// its results out after its traversal kernel, on the chunk's lane stream, so the transfers of one chunk overlap the kernels
// of the other lane (pinned buffers from cc_host_alloc make the copies truly asynchronous; pageable ones still work).
struct HostIO {
    const double* in = nullptr;
    uint32_t* verdict = nullptr;
    int32_t* fe = nullptr;
    int32_t* fs = nullptr;
    uint32_t* mask = nullptr;
    double* fk = nullptr;
};

// Everything of ccdev::Params that depends on the query and the plan only (io pointers and scratch are set per chunk).
void init_params(cc_query* Q, const Plan& pl, bool use_hoisted, int flags, ccdev::Params& P)
{
    cc_scene* S = Q->scene;
    memset(&P, 0, sizeof(P));
    P.nodes = S->d_nodes.p; P.tris = S->d_tris.p;
    P.slot_root = pl.d_slot_root.p; P.st_xf = pl.d_st_xf.p; P.st_aabb = pl.d_st_aabb.p; P.link_box = pl.d_link_box.p;
    P.emit_index = pl.d_emit.p;
    P.pairs = pl.d_pairs.p; P.pair_orig = pl.d_orig.p;
    P.steps = pl.d_steps.p;
    P.hoisted_mask = use_hoisted ? Q->d_hoisted_mask.p : nullptr;
    P.nq = Q->nq; P.ndyn = pl.ndyn; P.nstatic = pl.nstatic; P.npairs_dev = (int)pl.pairs.size(); P.nenv_dev = pl.nenv;
    P.nenv_total = Q->nenv_total; P.nsteps = (int)pl.steps.size(); P.ncols = Q->ncols; P.npose = pl.npose;
    P.hoisted_first_env = use_hoisted ? Q->hoisted_first_env : ccdev::kNone;
    P.hoisted_first_self = use_hoisted ? Q->hoisted_first_self : ccdev::kNone;
    P.mask_words = Q->mask_words;
    P.flags = flags | (S->counters ? ccdev::kFlagCounters : 0);
    memcpy(P.base, Q->base, sizeof(P.base));
    memcpy(P.center, Q->center, sizeof(P.center));
    P.affine_mask = Q->affine_mask; P.affine_col = Q->ncols - Q->naffine;
    memcpy(P.affine_axis, Q->affine_axis, sizeof(P.affine_axis)); memcpy(P.affine_qstart, Q->affine_qstart, sizeof(P.affine_qstart));
    P.tri_eps = S->tri_eps;
    P.err_flag = S->d_err;
    P.stats = S->d_stats.p;
    P.spill_cap = 2048;
    P.split = 1;
}

// per-device record of the kernels whose dynamic shared-memory limit has been raised
struct SmemOpted { bool fk2 = false, broad = false; int dist_blocks_per_sm = 0; };
SmemOpted g_smem_opted[64];

// FK of m configurations on stream cs: the two-phase kernel (cc_fk2_kernel) unless its per-block pose table does not fit in shared
// memory or CC_TUNE_FK2=0 asks for the one-thread-per-configuration form (A/B: profiles/r02y_ab_fk2.log)
int launch_fk(cc_scene* S, const ccdev::Params& P, int m, cudaStream_t cs)
{
    // the two-phase form pays when most links are rigidly attached to a few moving ones (PR2, arm + torso active: 8 varying steps,
    // 33 constant ones: 95.8 -> 103.0 M configs/s); a chain of varying steps only (HIRO-NX, every DOF active) runs its serial phase
    // with a quarter of the block and loses 7 % (profiles/r02y_ab_fk2.log).  CC_TUNE_FK2 = 0 / 1 forces one form.
    static const int tune_fk2 = getenv("CC_TUNE_FK2") ? atoi(getenv("CC_TUNE_FK2")) : -1;
    const int nvary = P.npose - 1, nconst = P.nsteps - nvary;
    const bool want_fk2 = tune_fk2 >= 0 ? tune_fk2 != 0 : ((P.flags & ccdev::kFlagPoseMode) != 0 || nconst >= 2 * nvary);
    const size_t fixed = ccdev::fk2_fixed_bytes(P.nq, P.nsteps, P.ndyn);
    const size_t row = (size_t)ccdev::fk2_row_doubles(P.npose) * sizeof(double);
    int group = 0;
    if (want_fk2)
        for (int g = ccdev::kFkThreads; g >= 32; g -= 32)
            if (fixed + g * row <= (g == 32 ? 200u : 48u) * 1024u) { group = g; break; }       // <= 48 KB per block where possible (4 - 5 resident blocks)
    if (group > 0) {
        const size_t smem = fixed + group * row;
        if (!g_smem_opted[S->device & 63].fk2) {
            // the opt-in is a property of the kernel on a device, shared by every scene of the process: raise it once to the most any
            // launch may ask for (a per-scene record of "what I asked for last" lets one scene lower another's limit)
            CC_CUDA(cudaFuncSetAttribute(ccdev::cc_fk2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            g_smem_opted[S->device & 63].fk2 = true;
        }
        ccdev::cc_fk2_kernel<<<(m + group - 1) / group, ccdev::kFkThreads, smem, cs>>>(P, group);
    } else {
        const size_t fk_smem = ccdev::fk_smem_bytes(P.nq, P.nsteps, P.ndyn);
        if (fk_smem > 227 * 1024) return fail(CC_ERR_UNSUPPORTED, "kinematic chain + environment tables exceed one block's shared memory");
        if (fk_smem > 48 * 1024)
            CC_CUDA(cudaFuncSetAttribute(ccdev::cc_fk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fk_smem));
        ccdev::cc_fk_kernel<<<(m + ccdev::kFkThreads - 1) / ccdev::kFkThreads, ccdev::kFkThreads, fk_smem, cs>>>(P);
    }
    CC_CUDA(cudaGetLastError());
    S->stats.launches++;
    return CC_OK;
}

int launch(cc_query* Q, const Plan& pl, bool use_hoisted, const double* d_q, int n, int flags, uint32_t* d_verdict,
           int32_t* d_fe, int32_t* d_fs, uint32_t* d_mask, double* d_fk, cudaStream_t st, const HostIO* hio = nullptr)
{
    cc_scene* S = Q->scene;
    if (n <= 0) return CC_OK;
    ccdev::Params P;
    init_params(Q, pl, use_hoisted, flags, P);
    const bool fk_only = (flags & ccdev::kFlagFkOnly) != 0;
    const int row_doubles = (flags & ccdev::kFlagPoseMode) ? Q->nq * 8 : Q->ncols;

    // CC_TUNE_EXTRA_SMEM (bytes per block) is a profiling knob: it trades resident traversal blocks for L1 capacity
    static const size_t extra_smem = getenv("CC_TUNE_EXTRA_SMEM") ? (size_t)atoi(getenv("CC_TUNE_EXTRA_SMEM")) : 0;
    const size_t tr_smem = ccdev::kWarpSmemBytes * ccdev::kWarpsPerBlock + extra_smem;
    if (S->trav_blocks_per_sm == 0) {
        const void* variants[6] = {(const void*)ccdev::cc_traverse_kernel<false, false>, (const void*)ccdev::cc_traverse_kernel<true, false>,
                                   (const void*)ccdev::cc_traverse_kernel<false, true>, (const void*)ccdev::cc_traverse_kernel<true, true>,
                                   (const void*)ccdev::cc_traverse_kernel<false, false, true>,
                                   (const void*)ccdev::cc_traverse_kernel<false, false, false, true>};
        int per_sm = 1 << 30;
        for (const void* kfn : variants) {
            CC_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tr_smem));
            int v = 0;
            CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kfn, ccdev::kWarpsPerBlock * 32, tr_smem));
            if (v < 1) return fail(CC_ERR_CUDA, "traversal kernel does not fit on an SM");
            per_sm = std::min(per_sm, v);
            if (getenv("CC_TUNE_CARVEOUT")) CC_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("CC_TUNE_CARVEOUT"))));
        }
        if (getenv("CC_TUNE_BLOCKS_PER_SM")) per_sm = std::min(per_sm, std::max(1, atoi(getenv("CC_TUNE_BLOCKS_PER_SM"))));
        S->trav_blocks_per_sm = per_sm;
    }
    // spill areas: one per pipeline lane, sized for a full persistent grid
    const int full_warps = S->trav_blocks_per_sm * S->num_sms * ccdev::kWarpsPerBlock;
    // small batches: spread every configuration over 2, 4 or 8 warps (as many as keep the work items within ~2 waves and
    // leave each part at least one block of 32 pairs); CC_TUNE_SPLIT forces a value (1 = off)
    int split = 1;
    if (!fk_only) {
        static const int forced = getenv("CC_TUNE_SPLIT") ? atoi(getenv("CC_TUNE_SPLIT")) : 0;
        const int by_load = (int)std::min<long long>(8, (2LL * full_warps) / std::max(1, n));
        const int by_pairs = ((int)pl.pairs.size() + 31) / 32;
        int want = forced > 0 ? forced : std::min(by_load, by_pairs);
        while (split * 2 <= want && split < 8) split *= 2;
    }
    const bool contacts = Q->contact_out != nullptr && (flags & CC_BATCH_ALL_PAIRS);
    if (contacts) split = 1;
    // ---- chunking.  Big batches are cut into chunks that alternate between two internal streams: while the persistent
    // traversal kernel of chunk i drains its last (long) configurations, the kernels of chunk i+1 fill the SMs it has already
    // left, and (host-buffer entry) the copies of one chunk overlap the kernels of the other lane.  Chunks of 256k: with 128k
    // the per-chunk tails cost 10 % on the PR2 workload (profiles/r02a_knobs.log).  Scratch per set stays below 1 GB.
    const size_t per_cfg = std::max<size_t>(64, scratch_bytes_per_config(pl));
    int chunk = (int)std::min<size_t>(262144, std::max<size_t>(1024, ((size_t)1 << 30) / per_cfg));
    static const int tune_chunk = getenv("CC_TUNE_CHUNK") ? atoi(getenv("CC_TUNE_CHUNK")) : 0;      // profiling knob: configurations per chunk
    if (tune_chunk > 0) chunk = (tune_chunk + 127) / 128 * 128;
    else if (n >= 98304 && n <= chunk) chunk = ((n + 1) / 2 + 127) / 128 * 128;      // a single chunk overlaps nothing: two halves (+2.5 % on
                                                                                      // the 100k-configuration WAM batch, profiles/r02r_chunk_sweep.log)
    chunk = std::min(chunk, (n + 127) / 128 * 128);
    // host-buffer batches of three chunks or more start with a quarter and a half chunk: the first kernels wait for their
    // slice's H2D copy (16.8 MB per full PR2 chunk), a short first slice gets the device going sooner (CC_TUNE_RAMP=0: off)
    static const bool tune_ramp = !(getenv("CC_TUNE_RAMP") && atoi(getenv("CC_TUNE_RAMP")) == 0);
    std::vector<int> sizes;
    {
        const bool ramp = tune_ramp && hio && hio->in && n >= 3 * chunk && !fk_only;
        int left = n;
        if (ramp) { sizes.push_back(chunk / 4 / 128 * 128); sizes.push_back(chunk / 2 / 128 * 128); left -= sizes[0] + sizes[1]; }
        while (left > 0) { sizes.push_back(std::min(chunk, left)); left -= sizes.back(); }
    }
    const int nchunks = (int)sizes.size();
    const bool overlap = nchunks >= 2;
    static const int nlanes = std::min(kMaxLanes, std::max(2, getenv("CC_TUNE_LANES") ? atoi(getenv("CC_TUNE_LANES")) : 4));     // pipeline lanes: 4 > 3 > 2 by about 1 % (profiles/r02n_ab.log)
    static const int tune_skip_fk = getenv("CC_TUNE_SKIP_FK") ? atoi(getenv("CC_TUNE_SKIP_FK")) : 0;  // profiling knob: FK only for the first
                                                                                                     // chunk of each lane (timing without FK)
    const bool timeline = S->trace_timeline && !fk_only;
    if (timeline) {
        while ((int)S->tl_events.size() < 3 * nchunks) { cudaEvent_t e; CC_CUDA(cudaEventCreate(&e)); S->tl_events.push_back(e); }
        S->tl_lane.assign(nchunks, 0);
    }
    {
        int rc = ensure_scratch(Q, pl, chunk, overlap ? nlanes : 1);
        if (rc != CC_OK) return rc;
    }
    P.cap = chunk;
    if (overlap && !S->lane_stream[0]) {
        for (int i = 0; i < kMaxLanes; ++i) {
            CC_CUDA(cudaStreamCreateWithFlags(&S->lane_stream[i], cudaStreamNonBlocking));
            CC_CUDA(cudaEventCreateWithFlags(&S->ev_join[i], cudaEventDisableTiming));
        }
        CC_CUDA(cudaEventCreateWithFlags(&S->ev_fork, cudaEventDisableTiming));
    }

    if (Q->trace_work && !fk_only) { CC_CUDA(Q->d_work.reserve((size_t)n)); Q->work_n = n; }
    if (contacts) { P.contact_out = Q->contact_out; P.contact_count = Q->contact_count; P.contact_cap = Q->contact_cap; }
    const bool use_tables = Q->ntables > 0 && &pl == &Q->mainl && !(flags & ccdev::kFlagPoseMode) && !fk_only;
    if (use_tables) { P.pair_table = Q->d_pair_table.p; P.tables = Q->d_tables.p; P.table_bits = Q->d_table_bits.p; }
    // full batches: the broadphase (and the certificate lookups) run one thread per configuration in cc_broad_kernel and hand the
    // traversal one bit per pair; CC_TUNE_CAND=0 keeps the broadphase inside the traversal warp (A/B: profiles/r02k_ab_cand.log)
    static const bool tune_cand = !(getenv("CC_TUNE_CAND") && atoi(getenv("CC_TUNE_CAND")) == 0);
    const bool use_cand = tune_cand && split == 1 && !contacts && !fk_only && !pl.pairs.empty();
    P.cand_words = (int)((pl.pairs.size() + 31) / 32);
    const Plan::Split* sp = split > 1 ? &pl.split[split == 2 ? 0 : (split == 4 ? 1 : 2)] : nullptr;
    if (sp && use_tables) P.pair_table = Q->d_pair_table_split[split == 2 ? 0 : (split == 4 ? 1 : 2)].p;
    if (sp) {
        P.pairs = sp->d_pairs.p; P.pair_orig = sp->d_orig.p; P.part_begin = sp->d_part_begin.p; P.part_nenv = sp->d_part_nenv.p;
        P.split = split;
        P.flags |= ccdev::kFlagInitOutputs;
    } else P.split = 1;
    if (!fk_only && (S->spill_warps < full_warps || (overlap && !S->d_spill[nlanes - 1].p))) {
        CC_CUDA(cudaDeviceSynchronize());            // (first launches only) nobody may still be using the old areas
        for (int k = 0; k < (overlap ? nlanes : 1); ++k) CC_CUDA(S->d_spill[k].reserve((size_t)full_warps * P.spill_cap));
        S->spill_warps = full_warps;
    }
    CC_CUDA(cudaEventRecord(S->ev0, st));
    if (overlap) {
        CC_CUDA(cudaEventRecord(S->ev_fork, st));
        for (int i = 0; i < nlanes; ++i) CC_CUDA(cudaStreamWaitEvent(S->lane_stream[i], S->ev_fork, 0));
    }
    // results of chunk i go back to the host only after chunk i+1 has been enqueued: a copy into pageable memory blocks
    // the calling thread, and the device must already have the next chunk's work queued by then
    int pend_off = -1, pend_m = 0;
    cudaStream_t pend_cs = nullptr;
    auto copy_out = [&](int off, int m, cudaStream_t cs) -> cudaError_t {
        cudaError_t e = cudaSuccess;
        if (hio->verdict && e == cudaSuccess) e = cudaMemcpyAsync(hio->verdict + off, d_verdict + off, (size_t)m * 4, cudaMemcpyDeviceToHost, cs);
        if (hio->fe && e == cudaSuccess) e = cudaMemcpyAsync(hio->fe + off, d_fe + off, (size_t)m * 4, cudaMemcpyDeviceToHost, cs);
        if (hio->fs && e == cudaSuccess) e = cudaMemcpyAsync(hio->fs + off, d_fs + off, (size_t)m * 4, cudaMemcpyDeviceToHost, cs);
        if (hio->mask && e == cudaSuccess)
            e = cudaMemcpyAsync(hio->mask + (size_t)off * Q->mask_words, d_mask + (size_t)off * Q->mask_words,
                                (size_t)m * Q->mask_words * 4, cudaMemcpyDeviceToHost, cs);
        if (hio->fk && e == cudaSuccess)
            e = cudaMemcpyAsync(hio->fk + (size_t)off * Q->nq * 8, d_fk + (size_t)off * Q->nq * 8, (size_t)m * Q->nq * 64,
                                cudaMemcpyDeviceToHost, cs);
        return e;
    };
    int off = 0;
    for (int ci = 0; ci < nchunks; off += sizes[ci], ++ci) {
        const int m = sizes[ci];
        const int k = overlap ? (ci % nlanes) : 0;
        cudaStream_t cs = overlap ? S->lane_stream[k] : st;
        if (hio && hio->in && row_doubles > 0)
            CC_CUDA(cudaMemcpyAsync(const_cast<double*>(d_q) + (size_t)off * row_doubles, hio->in + (size_t)off * row_doubles,
                                    (size_t)m * row_doubles * sizeof(double), cudaMemcpyHostToDevice, cs));
        P.n = m;
        P.q = d_q ? d_q + (size_t)off * row_doubles : nullptr;
        P.verdict = d_verdict ? d_verdict + off : nullptr;
        P.first_env = d_fe ? d_fe + off : nullptr;
        P.first_self = d_fs ? d_fs + off : nullptr;
        P.pair_mask = d_mask ? d_mask + (size_t)off * Q->mask_words : nullptr;
        P.fk_out = d_fk ? d_fk + (size_t)off * Q->nq * 8 : nullptr;
        P.work_out = (Q->trace_work && !fk_only) ? Q->d_work.p + off : nullptr;
        P.xf32 = Q->d_xf32[k].p; P.aabb = Q->d_aabb[k].p; P.fksave = Q->d_fksave[k].p;
        P.counters = Q->d_counters[k].p;
        P.cand = Q->d_cand[k].p;
        if (!fk_only) CC_CUDA(cudaMemsetAsync(Q->d_counters[k].p, 0, sizeof(int32_t), cs));
        if (timeline) { S->tl_lane[ci] = k; CC_CUDA(cudaEventRecord(S->tl_events[3 * ci], cs)); }
        if (!(tune_skip_fk && ci >= 2 && !fk_only)) {
            const int rc = launch_fk(S, P, m, cs);
            if (rc != CC_OK) return rc;
        }
        if (use_cand) {
            const int kb = (P.flags & ccdev::kFlagEnv) ? 0 : P.nenv_dev, ke = (P.flags & ccdev::kFlagSelf) ? P.npairs_dev : P.nenv_dev;
            const int bwords = std::max(0, ((ke + 31) >> 5) - (kb >> 5));
            if (bwords > 0) {
                const int rounds = (bwords + ccdev::kBroadMaxWarps - 1) / ccdev::kBroadMaxWarps;
                const int bw = (bwords + rounds - 1) / rounds;          // warps per block: the words spread evenly over the rounds
                const int group = std::max(1, std::min(32, (96 * 1024) / std::max(1, P.ndyn * 16)));     // configurations per block
                const size_t bsmem = (size_t)group * std::max(1, P.ndyn) * 16;
                if (!g_smem_opted[S->device & 63].broad) {
                    CC_CUDA(cudaFuncSetAttribute(ccdev::cc_broad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
                    g_smem_opted[S->device & 63].broad = true;
                }
                ccdev::cc_broad_kernel<<<(m + group - 1) / group, 32 * bw, bsmem, cs>>>(P, group);
                CC_CUDA(cudaGetLastError());
                S->stats.launches++;
            }
        }
        if (timeline) CC_CUDA(cudaEventRecord(S->tl_events[3 * ci + 1], cs));
        if (fk_only) {
            if (hio) CC_CUDA(copy_out(off, m, cs));
            continue;
        }
        // persistent traversal grid: a whole number of waves, never more warps than configurations
        int grid = S->trav_blocks_per_sm * S->num_sms;
        const int need = (int)(((long long)m * split + ccdev::kWarpsPerBlock - 1) / ccdev::kWarpsPerBlock);
        if (grid > need) grid = need;
        P.spill = S->d_spill[k].p;
        if (contacts) ccdev::cc_traverse_kernel<false, false, true><<<grid, ccdev::kWarpsPerBlock * 32, tr_smem, cs>>>(P);
        else if (use_cand) ccdev::cc_traverse_kernel<false, false, false, true><<<grid, ccdev::kWarpsPerBlock * 32, tr_smem, cs>>>(P);
        else if (use_tables) {
            if (split > 1) ccdev::cc_traverse_kernel<true, true><<<grid, ccdev::kWarpsPerBlock * 32, tr_smem, cs>>>(P);
            else ccdev::cc_traverse_kernel<false, true><<<grid, ccdev::kWarpsPerBlock * 32, tr_smem, cs>>>(P);
        } else {
            if (split > 1) ccdev::cc_traverse_kernel<true, false><<<grid, ccdev::kWarpsPerBlock * 32, tr_smem, cs>>>(P);
            else ccdev::cc_traverse_kernel<false, false><<<grid, ccdev::kWarpsPerBlock * 32, tr_smem, cs>>>(P);
        }
        CC_CUDA(cudaGetLastError());
        S->stats.launches++;
        if (!pl.apairs.empty()) {
            // analytic link pairs on the exact shapes, merged into what the traversal has just written
            ccdev::AnalyticParams AP;
            AP.pairs = pl.d_apairs.p; AP.orig = pl.d_aorig.p; AP.npairs = (int)pl.apairs.size(); AP.nenv = pl.anenv;
            AP.slot_prim = pl.d_slot_prim.p; AP.prims = pl.d_prims.p;
            const long long items = (long long)m * AP.npairs;
            ccdev::cc_analytic_kernel<<<(unsigned)((items + 127) / 128), 128, 0, cs>>>(P, AP);
            CC_CUDA(cudaGetLastError());
            S->stats.launches++;
        }
        if (timeline) CC_CUDA(cudaEventRecord(S->tl_events[3 * ci + 2], cs));
        if (hio) {
            if (pend_off >= 0) CC_CUDA(copy_out(pend_off, pend_m, pend_cs));
            pend_off = off; pend_m = m; pend_cs = cs;
        }
    }
    if (hio && pend_off >= 0) CC_CUDA(copy_out(pend_off, pend_m, pend_cs));
    if (overlap) {
        for (int i = 0; i < nlanes; ++i) {
            CC_CUDA(cudaEventRecord(S->ev_join[i], S->lane_stream[i]));
            CC_CUDA(cudaStreamWaitEvent(st, S->ev_join[i], 0));
        }
    }
    CC_CUDA(cudaEventRecord(S->ev1, st));
    S->ev_pending = true;
    if (!fk_only) S->stats.configs += (uint64_t)n;
    return CC_OK;
}

// Minimum separation of n configurations (host buffers in and out, one stream, synchronous): FK kernel, then the
// branch-and-bound distance kernel of cc_distance.cuh, chunk by chunk.
int launch_distance(cc_query* Q, const Plan& pl, bool pose_mode, const double* h_in, int n, int flags, double max_distance,
                    double rel_err, double abs_err, double* h_dist, int32_t* h_pair, double* h_pts, bool per_pair = false, double bound = -1.0)
{
    cc_scene* S = Q->scene;
    if (n <= 0) return CC_OK;
    CC_CUDA(cudaSetDevice(S->device));
    ccdev::Params P;
    init_params(Q, pl, false, (flags & (CC_BATCH_ENV | CC_BATCH_SELF)) | (pose_mode ? ccdev::kFlagPoseMode : 0), P);
    const int row_doubles = pose_mode ? Q->nq * 8 : Q->ncols;
    const size_t per_cfg = std::max<size_t>(64, scratch_bytes_per_config(pl));
    int chunk = (int)std::min<size_t>(65536, std::max<size_t>(1024, ((size_t)1 << 29) / per_cfg));
    const int outs_per_cfg = per_pair ? std::max(1, (int)pl.pairs.size()) : 1;       // per-pair mode: one result per (configuration, pair)
    if (per_pair) chunk = std::max(1, std::min(chunk, (1 << 22) / outs_per_cfg));
    chunk = std::min(chunk, (n + 127) / 128 * 128);
    int rc = ensure_scratch(Q, pl, chunk, 1);
    if (rc != CC_OK) return rc;
    P.cap = chunk;
    CC_CUDA(Q->d_q.reserve(std::max<size_t>(1, (size_t)chunk * row_doubles)));
    CC_CUDA(Q->d_dist.reserve((size_t)chunk * outs_per_cfg)); CC_CUDA(Q->d_dpair.reserve((size_t)chunk * outs_per_cfg));
    CC_CUDA(Q->d_dpts.reserve((size_t)chunk * 6));
    const size_t d_smem = ccdev::kDistWarpSmemBytes * ccdev::kWarpsPerBlock;
    int& dist_blocks_per_sm = g_smem_opted[S->device & 63].dist_blocks_per_sm;      // per device: the opt-in below is
    if (dist_blocks_per_sm == 0) {
        CC_CUDA(cudaFuncSetAttribute(ccdev::cc_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d_smem));
        CC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&dist_blocks_per_sm, ccdev::cc_distance_kernel, ccdev::kWarpsPerBlock * 32, d_smem));
        if (dist_blocks_per_sm < 1) { dist_blocks_per_sm = 0; return fail(CC_ERR_CUDA, "distance kernel does not fit on an SM"); }
    }
    ccdev::DistOut D;
    D.min_distance = Q->d_dist.p; D.min_pair = Q->d_dpair.p; D.points = h_pts ? Q->d_dpts.p : nullptr;
    D.dist_max = max_distance > 0 ? (float)max_distance : 0.f;
    D.bound = bound > 0 ? nextafterf((float)bound, 3.0e38f) : 0.f;      // never below the double it came from
    D.rel_err = rel_err > 0 ? (float)rel_err : 0.f;
    D.abs_err = abs_err > 0 ? (float)abs_err : 0.f;
    D.per_pair = per_pair ? 1 : 0;
    if (per_pair) { D.min_pair = nullptr; D.points = nullptr; }
    cudaStream_t st = 0;
    for (int off = 0; off < n; off += chunk) {
        const int m = std::min(chunk, n - off);
        if (row_doubles > 0)
            CC_CUDA(cudaMemcpyAsync(Q->d_q.p, h_in + (size_t)off * row_doubles, (size_t)m * row_doubles * sizeof(double), cudaMemcpyHostToDevice, st));
        P.n = m;
        P.q = Q->d_q.p;
        P.xf32 = Q->d_xf32[0].p; P.aabb = Q->d_aabb[0].p; P.fksave = Q->d_fksave[0].p; P.counters = Q->d_counters[0].p;
        CC_CUDA(cudaMemsetAsync(Q->d_counters[0].p, 0, sizeof(int32_t), st));
        {
            const int rc2 = launch_fk(S, P, m, st);
            if (rc2 != CC_OK) return rc2;
        }
        int grid = dist_blocks_per_sm * S->num_sms;
        const long long need = ((long long)m * outs_per_cfg + ccdev::kWarpsPerBlock - 1) / ccdev::kWarpsPerBlock;
        if (grid > need) grid = (int)need;
        ccdev::cc_distance_kernel<<<grid, ccdev::kWarpsPerBlock * 32, d_smem, st>>>(P, D);
        CC_CUDA(cudaGetLastError());
        S->stats.launches += 1;
        CC_CUDA(cudaMemcpyAsync(h_dist + (size_t)off * outs_per_cfg, Q->d_dist.p, (size_t)m * outs_per_cfg * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (h_pair && !per_pair) CC_CUDA(cudaMemcpyAsync(h_pair + off, Q->d_dpair.p, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (h_pts && !per_pair) CC_CUDA(cudaMemcpyAsync(h_pts + (size_t)off * 6, Q->d_dpts.p, (size_t)m * 6 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CC_CUDA(cudaStreamSynchronize(st));
    }
    return CC_OK;
}

int check_err_flag(cc_scene* S)
{
    // callers have synchronised the stream(s) that could have set it
    const int32_t h = *(volatile int32_t*)S->h_err;
    if (h) {
        *(volatile int32_t*)S->h_err = 0;
        return fail(CC_ERR_NOMEM, "device traversal stack overflowed its spill area");
    }
    return CC_OK;
}

HBody* get_body(cc_scene* s, int body)
{
    if (!s || body < 0 || body >= (int)s->bodies.size() || s->bodies[body]->removed) return nullptr;
    return s->bodies[body];
}

// mirror a mutating call onto the peer replicas of a multi-device scene
template <typename F>
int forward_peers(cc_scene* s, F f)
{
    for (cc_scene* p : s->peers) { const int rc = f(p); if (rc != CC_OK) return rc; }
    return CC_OK;
}

}  // namespace


#include "cc_certify.inl"

// ================================================================================================ C ABI
extern "C" {

// used by the other translation units of this library (cc_planner.cpp) so that cc_last_error stays one per thread
void cc_internal_set_error(const char* msg) { g_err = msg ? msg : ""; }

const char* cc_last_error(void) { return g_err.c_str(); }
int cc_version(void) { return 100; }

int cc_scene_create(int cuda_device, cc_scene** out)
{
    if (!out) return fail(CC_ERR_INVALID, "out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(CC_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
    if (cuda_device < 0 || cuda_device >= ndev) return fail(CC_ERR_INVALID, "bad device index");
    CC_CUDA(cudaSetDevice(cuda_device));
    cc_scene* s = new cc_scene();
    s->device = cuda_device;
    cudaDeviceProp prop;
    CC_CUDA(cudaGetDeviceProperties(&prop, cuda_device));
    s->num_sms = prop.multiProcessorCount;
    CC_CUDA(cudaHostAlloc((void**)&s->h_err, sizeof(int32_t), cudaHostAllocMapped | cudaHostAllocPortable));
    *s->h_err = 0;
    CC_CUDA(cudaHostGetDevicePointer((void**)&s->d_err, s->h_err, 0));
    CC_CUDA(s->d_stats.reserve(2));
    CC_CUDA(cudaMemset(s->d_stats.p, 0, 2 * sizeof(unsigned long long)));
    CC_CUDA(cudaEventCreate(&s->ev0));
    CC_CUDA(cudaEventCreate(&s->ev1));
    *out = s;
    return CC_OK;
}

int cc_scene_create_multi(const int* cuda_devices, int ndevices, cc_scene** out)
{
    if (!out || !cuda_devices || ndevices < 1) return fail(CC_ERR_INVALID, "bad arguments");
    cc_scene* head = nullptr;
    int rc = cc_scene_create(cuda_devices[0], &head);
    if (rc != CC_OK) return rc;
    for (int i = 1; i < ndevices; ++i) {
        cc_scene* p = nullptr;
        rc = cc_scene_create(cuda_devices[i], &p);
        if (rc != CC_OK) { const std::string keep = g_err; cc_scene_destroy(head); g_err = keep; return rc; }
        head->peers.push_back(p);
    }
    cudaSetDevice(head->device);
    *out = head;
    return CC_OK;
}

int cc_scene_get_devices(const cc_scene* s, int* out_devices, int max_devices, int* out_count)
{
    if (!s || !out_count || max_devices < 0 || (max_devices > 0 && !out_devices)) return fail(CC_ERR_INVALID, "bad arguments");
    *out_count = 1 + (int)s->peers.size();
    for (int i = 0; i < *out_count && i < max_devices; ++i) out_devices[i] = i == 0 ? s->device : s->peers[i - 1]->device;
    return CC_OK;
}

int cc_scene_destroy(cc_scene* s)
{
    if (!s) return CC_OK;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    delete s;               // also tears the peer replicas of a multi-device scene down
    return CC_OK;
}

int cc_scene_add_body(cc_scene* s, const char* name, int nlinks, int is_robot, int* out_body)
{
    if (!s || nlinks < 0 || nlinks > 0xffff) return fail(CC_ERR_INVALID, "bad arguments");
    HBody* b = new HBody();
    b->name = name ? name : "";
    b->is_robot = is_robot != 0;
    b->links.resize(nlinks);
    s->bodies.push_back(b);
    int id = (int)s->bodies.size() - 1;
    b->attached.push_back(id);
    if (out_body) *out_body = id;
    return forward_peers(s, [&](cc_scene* p) { return cc_scene_add_body(p, name, nlinks, is_robot, nullptr); });
}

int cc_scene_remove_body(cc_scene* s, int body)
{
    HBody* b = get_body(s, body);
    if (!b) return fail(CC_ERR_INVALID, "bad body");
    b->removed = true;
    for (HLink& l : b->links) { l.verts.clear(); l.tris.clear(); l.nodes.clear(); l.tridata.clear(); }
    s->committed = false;
    return forward_peers(s, [&](cc_scene* p) { return cc_scene_remove_body(p, body); });
}

int cc_body_set_link_mesh(cc_scene* s, int body, int link, const char* link_name, const double* v, int nv,
                          const int32_t* tris, int ntris)
{
    HBody* b = get_body(s, body);
    if (!b || link < 0 || link >= (int)b->links.size()) return fail(CC_ERR_INVALID, "bad body/link");
    if (nv < 0 || ntris < 0 || (ntris > 0 && (!v || !tris))) return fail(CC_ERR_INVALID, "bad mesh");
    for (int i = 0; i < 3 * ntris; ++i)
        if (tris[i] < 0 || tris[i] >= nv) return fail(CC_ERR_INVALID, "triangle index out of range");
    HLink& L = b->links[link];
    L.name = link_name ? link_name : "";
    L.verts.assign(v, v + 3 * (size_t)nv);
    L.tris.assign(tris, tris + 3 * (size_t)ntris);
    L.dirty = true;
    s->committed = false;
    return forward_peers(s, [&](cc_scene* p) { return cc_body_set_link_mesh(p, body, link, link_name, v, nv, tris, ntris); });
}

int cc_body_set_link_primitives(cc_scene* s, int body, int link, const cc_geom* geoms, int ngeoms)
{
    HBody* b = get_body(s, body);
    if (!b || link < 0 || link >= (int)b->links.size() || ngeoms < 0 || (ngeoms > 0 && !geoms)) return fail(CC_ERR_INVALID, "bad arguments");
    for (int i = 0; i < ngeoms; ++i) {
        if (geoms[i].type < CC_GEOM_BOX || geoms[i].type > CC_GEOM_CAPSULE) return fail(CC_ERR_INVALID, "unknown primitive type");
        if (!(geoms[i].dims[0] > 0)) return fail(CC_ERR_INVALID, "primitive with a non-positive size");
    }
    HLink& L = b->links[link];
    L.prims.assign(geoms, geoms + ngeoms);
    L.dirty = true;              // the root box has to hold the exact shapes
    s->committed = false;
    return forward_peers(s, [&](cc_scene* p) { return cc_body_set_link_primitives(p, body, link, geoms, ngeoms); });
}

int cc_scene_set_analytic_primitives(cc_scene* s, int enabled)
{
    if (!s) return fail(CC_ERR_INVALID, "scene is NULL");
    s->analytic = enabled != 0;
    return forward_peers(s, [&](cc_scene* p) { return cc_scene_set_analytic_primitives(p, enabled); });
}

int cc_body_set_link_enabled(cc_scene* s, int body, const uint8_t* enabled)
{
    HBody* b = get_body(s, body);
    if (!b || !enabled) return fail(CC_ERR_INVALID, "bad arguments");
    for (size_t i = 0; i < b->links.size(); ++i) b->links[i].enabled = enabled[i] != 0;
    return forward_peers(s, [&](cc_scene* p) { return cc_body_set_link_enabled(p, body, enabled); });
}

int cc_body_set_link_transforms(cc_scene* s, int body, const double* T8)
{
    HBody* b = get_body(s, body);
    if (!b || !T8) return fail(CC_ERR_INVALID, "bad arguments");
    for (size_t i = 0; i < b->links.size(); ++i) memcpy(b->links[i].xf, T8 + 8 * i, 8 * sizeof(double));
    s->pose_epoch++;
    return forward_peers(s, [&](cc_scene* p) { return cc_body_set_link_transforms(p, body, T8); });
}

static int set_attached_one(cc_scene* s, int a, int bb, int attached);
int cc_body_set_attached(cc_scene* s, int a, int bb, int attached)
{
    if (!s) return fail(CC_ERR_INVALID, "scene is NULL");
    const int rc = set_attached_one(s, a, bb, attached);
    if (rc != CC_OK) return rc;
    return forward_peers(s, [&](cc_scene* p) { return set_attached_one(p, a, bb, attached); });
}

static int set_attached_one(cc_scene* s, int a, int bb, int attached)
{
    HBody* A = get_body(s, a); HBody* B = get_body(s, bb);
    if (!A || !B) return fail(CC_ERR_INVALID, "bad body");
    for (int pass = 0; pass < 2; ++pass) {
        HBody* x = pass ? B : A; int o = pass ? a : bb;
        auto it = std::find(x->attached.begin(), x->attached.end(), o);
        if (attached && it == x->attached.end()) x->attached.push_back(o);
        if (!attached && it != x->attached.end() && o != (pass ? bb : a)) x->attached.erase(it);
    }
    return CC_OK;
}

int cc_body_set_grabbed(cc_scene* s, int body, int grabbed_body, int link, const double* rel8, const int32_t* noncolliding,
                        int n_noncolliding)
{
    HBody* B = get_body(s, body); HBody* G = get_body(s, grabbed_body);
    if (!B || !G || body == grabbed_body || link < 0 || link >= (int)B->links.size() || !rel8 || n_noncolliding < 0 ||
        (n_noncolliding > 0 && !noncolliding))
        return fail(CC_ERR_INVALID, "bad arguments");
    for (int k = 0; k < n_noncolliding; ++k) {
        HBody* O = get_body(s, noncolliding[2 * k]);
        if (!O || noncolliding[2 * k + 1] < 0 || noncolliding[2 * k + 1] >= (int)O->links.size()) return fail(CC_ERR_INVALID, "bad non-colliding link");
    }
    HBody::Grab* g = nullptr;
    for (HBody::Grab& x : B->grabs) if (x.body == grabbed_body) g = &x;
    if (!g) { B->grabs.push_back(HBody::Grab()); g = &B->grabs.back(); }
    g->body = grabbed_body; g->link = link;
    g->rel.assign(rel8, rel8 + 8 * G->links.size());
    g->noncolliding.assign(noncolliding, noncolliding + 2 * (size_t)n_noncolliding);
    const int rc = set_attached_one(s, body, grabbed_body, 1);
    if (rc != CC_OK) return rc;
    return forward_peers(s, [&](cc_scene* p) { return cc_body_set_grabbed(p, body, grabbed_body, link, rel8, noncolliding, n_noncolliding); });
}

int cc_body_release_grabbed(cc_scene* s, int body, int grabbed_body)
{
    HBody* B = get_body(s, body);
    if (!B) return fail(CC_ERR_INVALID, "bad body");
    for (size_t i = 0; i < B->grabs.size();) {
        if (grabbed_body < 0 || B->grabs[i].body == grabbed_body) {
            if (get_body(s, B->grabs[i].body)) set_attached_one(s, body, B->grabs[i].body, 0);
            B->grabs.erase(B->grabs.begin() + i);
        } else ++i;
    }
    return forward_peers(s, [&](cc_scene* p) { return cc_body_release_grabbed(p, body, grabbed_body); });
}

int cc_body_set_kinematics(cc_scene* s, int body, int ndof, int njoints, const cc_joint* joints, const uint8_t* affects)
{
    HBody* b = get_body(s, body);
    if (!b || ndof < 0 || njoints < 0 || (njoints > 0 && !joints) || (ndof > 0 && !affects)) return fail(CC_ERR_INVALID, "bad arguments");
    const int nl = (int)b->links.size();
    for (int k = 0; k < njoints; ++k) {
        const cc_joint& j = joints[k];
        if (j.parent_link < 0 || j.parent_link >= nl || j.child_link < 0 || j.child_link >= nl) return fail(CC_ERR_INVALID, "joint link out of range");
        if (j.dof_index >= ndof || j.mimic_dof >= ndof) return fail(CC_ERR_INVALID, "joint dof out of range");
        if (j.mimic_dof < -1 && (j.mimic_dof != CC_MIMIC_EXTERNAL || j.reserved < 0 || j.reserved >= 4096 || j.dof_index >= 0))
            return fail(CC_ERR_INVALID, "externally evaluated mimic joint: mimic_dof = CC_MIMIC_EXTERNAL, reserved = its column ordinal, dof_index = -1");
        if (j.type != CC_JOINT_REVOLUTE && j.type != CC_JOINT_PRISMATIC) return fail(CC_ERR_UNSUPPORTED, "only 1-DOF revolute/prismatic joints");
    }
    b->ndof = ndof;
    b->joints.assign(joints, joints + njoints);
    b->affects.assign(affects, affects + (size_t)ndof * nl);
    if ((int)b->dofvals.size() != ndof) b->dofvals.assign(ndof, 0.0);
    b->mimic_prog.assign((size_t)njoints, ccexpr::Program());
    return forward_peers(s, [&](cc_scene* p) { return cc_body_set_kinematics(p, body, ndof, njoints, joints, affects); });
}

// ---- mimic equations the chain program cannot hold: evaluated on the host, handed to the device as columns of q
namespace {
// The value a mimic joint takes for the evaluated `ev` (kinbody.cpp:2326-2375; same rule as ccdev::joint_transform).
inline double mimic_limit_rule(double ev, const cc_joint& j)
{
    if (j.is_circular) return ev;
    if (ev < j.lower) { if (ev >= j.lower - CC_EPSILON_JOINT_LIMIT || ev < j.lower - CC_EPSILON_EVAL_JOINT_LIMIT) return j.lower; }
    else if (ev > j.upper) { if (ev <= j.upper + CC_EPSILON_JOINT_LIMIT || ev > j.upper + CC_EPSILON_EVAL_JOINT_LIMIT) return j.upper; }
    return ev;
}

// rows [active joint values | affine values] -> rows [active joint values | raw values of the CC_MIMIC_EXTERNAL joints by ordinal |
// affine values], the layout the batched entries take.  Per row this is the part of SetDOFValues' joint loop that produces joint
// VALUES (kinbody.cpp:2149-2224 clamp, 2286-2400 loop): DOF values clamped to their limits, joints in the listed (dependency)
// order, an equation reads DOF values directly and passive joints through the value the loop gave them.
int expand_rows_host(const std::vector<cc_joint>& J, int ndof, const double* dofvals, const std::vector<ccexpr::Program>& prog,
                     const int32_t* active, int nactive, int naffine, int nexternal, const double* q_in, int64_t n, double* q_out)
{
    const int nj = (int)J.size();
    std::vector<double> lo((size_t)ndof, 0.0), hi((size_t)ndof, 0.0);
    std::vector<uint8_t> circ((size_t)ndof, 1);
    for (const cc_joint& j : J)
        if (j.dof_index >= 0 && j.dof_index < ndof) { lo[j.dof_index] = j.lower; hi[j.dof_index] = j.upper; circ[j.dof_index] = j.is_circular != 0; }
    for (int k = 0; k < nj; ++k)
        if (J[k].mimic_dof == CC_MIMIC_EXTERNAL && !J[k].is_static && (k >= (int)prog.size() || prog[k].ops.empty())) {
            char msg[160];
            snprintf(msg, sizeof(msg), "joint %d is CC_MIMIC_EXTERNAL and has no equation (cc_body_set_mimic_equation)", k);
            return fail(CC_ERR_STATE, msg);
        }
    const int nin = nactive + naffine, nout = nactive + nexternal + naffine;
    // rows are independent: large batches are split over the host cores (a row costs a few hundred ns on one core, the device
    // checks one in 10); the first failing row, in row order, is the one reported
    struct Failure { int64_t row = -1; int joint = -1, err = 0; };
    // everything that does not depend on the row: the clamped values of the inactive DOFs, which joints read which column, and the
    // passive joints in the order the loop visits them
    std::vector<double> full0((size_t)ndof);
    for (int d = 0; d < ndof; ++d) full0[d] = circ[d] ? dofvals[d] : (dofvals[d] < lo[d] ? lo[d] : dofvals[d] > hi[d] ? hi[d] : dofvals[d]);
    std::vector<double> jv0((size_t)nj, 0.0);
    for (int k = 0; k < nj; ++k) if (J[k].dof_index >= 0) jv0[k] = full0[J[k].dof_index];
    std::vector<std::pair<int, int>> joint_of_col;                 // (joint, column) for joints that carry an active DOF
    for (int c = 0; c < nactive; ++c)
        for (int k = 0; k < nj; ++k) if (J[k].dof_index == active[c]) joint_of_col.push_back(std::make_pair(k, c));
    std::vector<int> passive;
    for (int k = 0; k < nj; ++k) if (!J[k].is_static && J[k].dof_index < 0) passive.push_back(k);
    auto run = [&](int64_t begin, int64_t end, Failure& f) {
    std::vector<double> full(full0), jv((size_t)nj);
    for (int64_t i = begin; i < end; ++i) {
        const double* in = q_in + (size_t)i * nin;
        double* out = q_out + (size_t)i * nout;
        for (int c = 0; c < nactive; ++c) {
            const int d = active[c];
            const double v = in[c];
            full[d] = circ[d] ? v : (v < lo[d] ? lo[d] : v > hi[d] ? hi[d] : v);
        }
        jv = jv0;
        for (const std::pair<int, int>& kc : joint_of_col) jv[kc.first] = full[active[kc.second]];
        for (int e = 0; e < nexternal; ++e) out[nactive + e] = 0.0;
        for (int k : passive) {
            const cc_joint& j = J[k];
            if (j.mimic_dof >= 0) jv[k] = mimic_limit_rule(j.mimic_a * full[j.mimic_dof] + j.mimic_b, j);
            else if (j.mimic_dof == CC_MIMIC_EXTERNAL) {
                int err = 0;
                const double ev = prog[k].eval(jv.data(), &err);
                if (err) { f.row = i; f.joint = k; f.err = err; return; }
                out[nactive + j.reserved] = ev;
                jv[k] = mimic_limit_rule(ev, j);
            } else jv[k] = j.is_circular ? 0.0 : (0.0 < j.lower ? j.lower : 0.0 > j.upper ? j.upper : 0.0);   // passive: 0 clamped (kinbody.cpp:2262-2275)
        }
        for (int c = 0; c < nactive; ++c) out[c] = in[c];
        for (int c = 0; c < naffine; ++c) out[nactive + nexternal + c] = in[nactive + c];
    }
    };
    const int64_t kRowsPerThread = 8192;
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(n / kRowsPerThread, 64), (int64_t)std::max(1u, std::thread::hardware_concurrency())));
    std::vector<Failure> failures((size_t)nthreads);
    if (nthreads == 1) run(0, n, failures[0]);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nthreads; ++t) pool.emplace_back(run, n * t / nthreads, n * (t + 1) / nthreads, std::ref(failures[t]));
        for (std::thread& th : pool) th.join();
    }
    for (const Failure& f : failures)             // thread t holds rows before those of thread t + 1
        if (f.row >= 0) {
            char msg[200];
            snprintf(msg, sizeof(msg), "the mimic equation of joint %d failed to evaluate at row %lld (error %d: %s)", f.joint, (long long)f.row, f.err,
                     f.err == 1 ? "division by zero" : f.err == 2 ? "sqrt of a negative value" : f.err == 3 ? "log of a non-positive value" : f.err == 4 ? "trigonometric domain" : "value not finite");
            return fail(CC_ERR_INVALID, msg);
        }
    return CC_OK;
}

int compile_equation(const char* equation, const char* const* joint_names, int njoints, ccexpr::Program& prog)
{
    int pos = 0;
    std::string why;
    if (!ccexpr::compile(equation, joint_names, njoints, prog, &pos, &why)) {
        char msg[400];
        snprintf(msg, sizeof(msg), "mimic equation '%.200s': %s at position %d", equation ? equation : "", why.c_str(), pos);
        return fail(CC_ERR_UNSUPPORTED, msg);
    }
    return CC_OK;
}
}   // namespace

int cc_body_set_mimic_equation(cc_scene* s, int body, int joint, const char* equation, const char* const* joint_names, int njoints)
{
    HBody* b = get_body(s, body);
    if (!b || !equation || !joint_names) return fail(CC_ERR_INVALID, "bad arguments");
    if (njoints != (int)b->joints.size()) return fail(CC_ERR_INVALID, "cc_body_set_mimic_equation: one name per joint of cc_body_set_kinematics, in the same order");
    if (joint < 0 || joint >= njoints || b->joints[joint].mimic_dof != CC_MIMIC_EXTERNAL)
        return fail(CC_ERR_INVALID, "cc_body_set_mimic_equation: the joint is not declared CC_MIMIC_EXTERNAL");
    ccexpr::Program prog;
    const int rc = compile_equation(equation, joint_names, njoints, prog);
    if (rc != CC_OK) return rc;
    // a variable must have its value when this joint's turn comes: a DOF, or a passive joint listed earlier
    for (int v : prog.vars)
        if (b->joints[v].dof_index < 0 && v >= joint) return fail(CC_ERR_INVALID, "cc_body_set_mimic_equation: the equation reads a passive joint listed after this one");
    b->mimic_prog.resize((size_t)njoints);
    b->mimic_prog[joint] = prog;
    return CC_OK;                      // host-side only: peer replicas take the expanded rows
}

int cc_query_expand_rows(cc_query* q, const double* q_in, int64_t n, double* q_out)
{
    if (!q || n < 0 || (n > 0 && (!q_in || !q_out))) return fail(CC_ERR_INVALID, "bad arguments");
    const HBody* B = q->scene->bodies[q->body];
    const int nactive = q->ncols - q->nexternal - q->naffine;
    std::vector<int32_t> active((size_t)nactive, -1);
    for (int d = 0; d < (int)q->col_of_dof.size(); ++d) if (q->col_of_dof[d] >= 0 && q->col_of_dof[d] < nactive) active[q->col_of_dof[d]] = d;
    return expand_rows_host(B->joints, B->ndof, B->dofvals.data(), B->mimic_prog, active.data(), nactive, q->naffine, q->nexternal, q_in, n, q_out);
}

int cc_debug_mimic_eval(const char* equation, const char* const* names, int nnames, const double* values, int64_t n, double* out, int32_t* err)
{
    if (!equation || nnames < 0 || (nnames && !names) || n < 0 || (n && (!out || (nnames && !values)))) return fail(CC_ERR_INVALID, "bad arguments");
    ccexpr::Program prog;
    const int rc = compile_equation(equation, names, nnames, prog);
    if (rc != CC_OK) return rc;
    for (int64_t i = 0; i < n; ++i) {
        int e = 0;
        out[i] = prog.eval(values + (size_t)i * nnames, &e);
        if (err) err[i] = e;
    }
    return CC_OK;
}

int cc_debug_expand_rows(const cc_joint* joints, int njoints, int ndof, const double* dofvals, const char* const* equations,
                         const char* const* joint_names, const int32_t* active, int nactive, int naffine, const double* q_in, int64_t n,
                         double* q_out)
{
    if (!joints || njoints <= 0 || ndof < 0 || (ndof && !dofvals) || !equations || !joint_names || nactive < 0 || (nactive && !active) ||
        naffine < 0 || n < 0 || (n && (!q_in || !q_out)))
        return fail(CC_ERR_INVALID, "bad arguments");
    std::vector<cc_joint> J(joints, joints + njoints);
    std::vector<ccexpr::Program> prog((size_t)njoints);
    int nexternal = 0;
    for (int k = 0; k < njoints; ++k) {
        if (J[k].dof_index >= ndof || J[k].mimic_dof >= ndof) return fail(CC_ERR_INVALID, "joint dof out of range");
        if (J[k].mimic_dof != CC_MIMIC_EXTERNAL) continue;
        if (J[k].reserved < 0 || J[k].reserved >= 4096) return fail(CC_ERR_INVALID, "external ordinal out of range");
        nexternal = std::max(nexternal, J[k].reserved + 1);
        if (!equations[k]) continue;
        const int rc = compile_equation(equations[k], joint_names, njoints, prog[k]);
        if (rc != CC_OK) return rc;
        for (int v : prog[k].vars)
            if (J[v].dof_index < 0 && v >= k) return fail(CC_ERR_INVALID, "the equation reads a passive joint listed after its own");
    }
    for (int c = 0; c < nactive; ++c) if (active[c] < 0 || active[c] >= ndof) return fail(CC_ERR_INVALID, "active DOF index out of range");
    return expand_rows_host(J, ndof, dofvals, prog, active, nactive, naffine, nexternal, q_in, n, q_out);
}

int cc_body_set_dof_values(cc_scene* s, int body, const double* values, int ndof)
{
    HBody* b = get_body(s, body);
    if (!b || ndof != b->ndof || (ndof > 0 && !values)) return fail(CC_ERR_INVALID, "bad arguments");
    b->dofvals.assign(values, values + ndof);
    s->pose_epoch++;
    return forward_peers(s, [&](cc_scene* p) { return cc_body_set_dof_values(p, body, values, ndof); });
}

int cc_body_set_nonadjacent_pairs(cc_scene* s, int body, const int32_t* packed, int n)
{
    HBody* b = get_body(s, body);
    if (!b || n < 0 || (n > 0 && !packed)) return fail(CC_ERR_INVALID, "bad arguments");
    const int nl = (int)b->links.size();
    for (int i = 0; i < n; ++i) {
        int l0 = packed[i] & 0xffff, l1 = (packed[i] >> 16) & 0xffff;
        if (l0 >= nl || l1 >= nl) return fail(CC_ERR_INVALID, "pair link out of range");
    }
    b->nonadjacent.assign(packed, packed + n);
    return forward_peers(s, [&](cc_scene* p) { return cc_body_set_nonadjacent_pairs(p, body, packed, n); });
}

int cc_scene_commit(cc_scene* s)
{
    if (!s) return fail(CC_ERR_INVALID, "scene is NULL");
    if (s->committed) return CC_OK;
    CC_CUDA(cudaSetDevice(s->device));
    std::vector<HNode> nodes;
    std::vector<float> tris;
    for (HBody* b : s->bodies) {
        if (b->removed) continue;
        for (HLink& L : b->links) {
            if (L.dirty) { build_link_bvh(L); expand_root_for_primitives(L); L.dirty = false; }
            if (L.nodes.empty()) { L.node_base = L.tri_base = -1; continue; }
            if ((nodes.size() & 1) == 0) nodes.push_back(HNode());   // odd base: sibling pairs (odd local index first) are 64-byte aligned
            L.node_base = (int)nodes.size();
            L.tri_base = (int)(tris.size() / 12);
            for (const HNode& n : L.nodes) {
                HNode g = n;
                if (g.left >= 0) g.left += L.node_base;
                else g.left = ~((~g.left) + L.tri_base);
                nodes.push_back(g);
            }
            tris.insert(tris.end(), L.tridata.begin(), L.tridata.end());
        }
    }
    if (nodes.size() >= ((size_t)1 << 26) || tris.size() / 12 >= ((size_t)1 << 26))
        return fail(CC_ERR_UNSUPPORTED, "scene exceeds 2^26 BVH nodes or triangles");
    s->total_nodes = (int)nodes.size();
    s->total_tris = (int)(tris.size() / 12);
    CC_CUDA(cudaDeviceSynchronize());    // no query may be in flight while the arrays are replaced
    CC_CUDA(s->d_nodes.reserve(nodes.size() * 2 + 2));
    CC_CUDA(s->d_tris.reserve(tris.size() / 4 + 3));
    if (!nodes.empty()) CC_CUDA(cudaMemcpy(s->d_nodes.p, nodes.data(), nodes.size() * sizeof(HNode), cudaMemcpyHostToDevice));
    if (!tris.empty()) CC_CUDA(cudaMemcpy(s->d_tris.p, tris.data(), tris.size() * sizeof(float), cudaMemcpyHostToDevice));
    s->committed = true;
    // peer replicas: the trees were just built here, hand them over instead of building them again per device
    for (cc_scene* p : s->peers) {
        if (p->bodies.size() != s->bodies.size()) return fail(CC_ERR_STATE, "replicas of a multi-device scene diverged");
        for (size_t b = 0; b < s->bodies.size(); ++b) {
            if (p->bodies[b]->links.size() != s->bodies[b]->links.size()) return fail(CC_ERR_STATE, "replicas of a multi-device scene diverged");
            for (size_t l = 0; l < s->bodies[b]->links.size(); ++l) {
                HLink& dst = p->bodies[b]->links[l];
                if (!dst.dirty) continue;
                dst.nodes = s->bodies[b]->links[l].nodes; dst.tridata = s->bodies[b]->links[l].tridata;
                dst.dirty = false;
            }
        }
        const int rc = cc_scene_commit(p);
        if (rc != CC_OK) return rc;
    }
    if (!s->peers.empty()) CC_CUDA(cudaSetDevice(s->device));
    return CC_OK;
}

// ---------------------------------------------------------------------------------------------- queries
// world AABB (base-centred) of a link's already inflated root box at its current pose, in double, rounded outwards
static void static_aabb(const HLink& L, const double center[3], float* o)
{
    for (int k = 0; k < 8; ++k) o[k] = 0.f;
    if (L.nodes.empty()) { o[0] = o[1] = o[2] = 1e30f; return; }          // never overlaps anything
    const double* x = L.xf;
    const double il = 2.0 / (x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
    const double q0 = x[0], q1 = x[1], q2 = x[2], q3 = x[3];
    const double Rm[9] = {1 - il * (q2 * q2 + q3 * q3), il * (q1 * q2 - q0 * q3), il * (q1 * q3 + q0 * q2),
                          il * (q1 * q2 + q0 * q3), 1 - il * (q1 * q1 + q3 * q3), il * (q2 * q3 - q0 * q1),
                          il * (q1 * q3 - q0 * q2), il * (q2 * q3 + q0 * q1), 1 - il * (q1 * q1 + q2 * q2)};
    for (int r = 0; r < 3; ++r) {
        double c = x[4 + r] - center[r], e = 0;
        for (int k = 0; k < 3; ++k) { c += Rm[3 * r + k] * (double)L.nodes[0].c[k]; e += (fabs(Rm[3 * r + k]) + 1e-6) * (double)L.nodes[0].e[k]; }
        const float cf = (float)c;
        e += fabs(c - (double)cf) + 1e-6;
        float ef = (float)e;
        if ((double)ef < e) ef = nextafterf(ef, std::numeric_limits<float>::infinity());
        o[r] = cf; o[4 + r] = ef;
    }
}

// ---- host mirror of the device's FP64 pose algebra (cc_device.cuh: same operation order, geometry.h:255-262, 430-480)
struct HXf { double q[4]; double t[3]; };

static void hxf_normalize(double* q)
{
    double f = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double eps = 2.220446049250313e-16;
    if (f < 1.0 - eps || f > 1.0 + eps) { f = sqrt(f); q[0] /= f; q[1] /= f; q[2] /= f; q[3] /= f; }
}

static HXf hxf_mul(const HXf& a, const HXf& b)
{
    HXf r;
    const double* q = a.q;
    const double xx = 2 * q[1] * q[1], xy = 2 * q[1] * q[2], xz = 2 * q[1] * q[3], xw = 2 * q[1] * q[0];
    const double yy = 2 * q[2] * q[2], yz = 2 * q[2] * q[3], yw = 2 * q[2] * q[0];
    const double zz = 2 * q[3] * q[3], zw = 2 * q[3] * q[0];
    r.t[0] = a.t[0] + ((1 - yy - zz) * b.t[0] + (xy - zw) * b.t[1] + (xz + yw) * b.t[2]);
    r.t[1] = a.t[1] + ((xy + zw) * b.t[0] + (1 - xx - zz) * b.t[1] + (yz - xw) * b.t[2]);
    r.t[2] = a.t[2] + ((xz - yw) * b.t[0] + (yz + xw) * b.t[1] + (1 - xx - yy) * b.t[2]);
    r.q[0] = a.q[0] * b.q[0] - a.q[1] * b.q[1] - a.q[2] * b.q[2] - a.q[3] * b.q[3];
    r.q[1] = a.q[0] * b.q[1] + a.q[1] * b.q[0] + a.q[2] * b.q[3] - a.q[3] * b.q[2];
    r.q[2] = a.q[0] * b.q[2] + a.q[2] * b.q[0] + a.q[3] * b.q[1] - a.q[1] * b.q[3];
    r.q[3] = a.q[0] * b.q[3] + a.q[3] * b.q[0] + a.q[1] * b.q[2] - a.q[2] * b.q[1];
    hxf_normalize(r.q);
    return r;
}

static HXf hxf_load(const double* p) { HXf r; for (int i = 0; i < 4; ++i) r.q[i] = p[i]; for (int i = 0; i < 3; ++i) r.t[i] = p[4 + i]; return r; }
static void hxf_store(const HXf& x, double* p) { for (int i = 0; i < 4; ++i) p[i] = x.q[i]; for (int i = 0; i < 3; ++i) p[4 + i] = x.t[i]; p[7] = 0.0; }

// L * J(value) * R of a joint: ccdev::joint_transform on the host.  qrow == NULL: a joint whose value does not depend on the rows
// of q (src_col < 0), the only use the product has for it; with a row it is the device-free restatement behind cc_debug_fk_program.
static HXf host_joint(const ccdev::FkStep& st, const double* qrow);
static HXf host_constant_joint(const ccdev::FkStep& st) { return host_joint(st, nullptr); }
static HXf host_joint(const ccdev::FkStep& st, const double* qrow)
{
    const HXf L = hxf_load(st.L);
    if (st.flags & 1) return L;
    double v = qrow && st.src_col >= 0 ? qrow[st.src_col] : st.src_const;
    if (!(st.flags & 8)) { if (v < st.src_lower) v = st.src_lower; else if (v > st.src_upper) v = st.src_upper; }
    if (st.flags & 2) {
        const double ev = st.a * v + st.b;
        v = ev;
        if (!(st.flags & 4)) {
            if (ev < st.lower) { if (ev >= st.lower - CC_EPSILON_JOINT_LIMIT) v = st.lower; else if (ev < st.lower - CC_EPSILON_EVAL_JOINT_LIMIT) v = st.lower; }
            else if (ev > st.upper) { if (ev <= st.upper + CC_EPSILON_JOINT_LIMIT) v = st.upper; else if (ev > st.upper + CC_EPSILON_EVAL_JOINT_LIMIT) v = st.upper; }
        }
    }
    HXf J; J.q[0] = 1; J.q[1] = J.q[2] = J.q[3] = 0; J.t[0] = J.t[1] = J.t[2] = 0;
    if (st.type == 1) {
        const double len = sqrt(st.axis[0] * st.axis[0] + st.axis[1] * st.axis[1] + st.axis[2] * st.axis[2]);
        if (len != 0) { const double sn = sin(v * 0.5) / len, cs = cos(v * 0.5); J.q[0] = cs; J.q[1] = st.axis[0] * sn; J.q[2] = st.axis[1] * sn; J.q[3] = st.axis[2] * sn; }
    } else { J.t[0] = st.axis[0] * v; J.t[1] = st.axis[1] * v; J.t[2] = st.axis[2] * v; }
    return hxf_mul(hxf_mul(L, J), hxf_load(st.R));
}

// Compile the chain program of a plan from the query's joint list (Q->steps: every joint, depth-first, with its current
// constant value): varying joints are walked on the device, everything rigidly attached to a varying joint's child (or to the
// base link) is folded into one constant per emitted link (ccdev::FkStep).  Re-run by cc_query_refresh when the values of
// inactive DOFs change.
static void compile_fk_program(const cc_query* Q, Plan& pl)
{
    const int nl = Q->nq;
    pl.steps.clear();
    pl.nsave = 0;
    pl.npose = 1;
    if (!Q->has_fk || Q->steps.empty()) return;
    static const bool fold = !(getenv("CC_TUNE_FK_FOLD") && atoi(getenv("CC_TUNE_FK_FOLD")) == 0);    // 0: walk every joint (A/B)
    std::vector<uint8_t> need(nl, 0);
    for (int i = (int)Q->steps.size() - 1; i >= 0; --i) {
        const ccdev::FkStep& st = Q->steps[i];
        if (pl.emit[st.child] >= 0) need[st.child] = 1;
        if (need[st.child]) need[st.parent] = 1;         // depth-first order: children come after their parent
    }
    // per link: the link whose pose it is rigidly attached to ("anchor": link 0 or the child of a varying joint), its pose in
    // that link's frame, and the constant steps waiting to be placed after the anchor's varying step
    std::vector<int> anchor(nl, 0);
    std::vector<HXf> rel(nl);
    std::vector<uint8_t> rel_identity(nl, 1);
    std::vector<std::vector<ccdev::FkStep>> consts(nl);      // indexed by anchor link
    std::vector<ccdev::FkStep> varying;                        // in depth-first order
    for (const ccdev::FkStep& st0 : Q->steps) {
        if (!need[st0.child]) continue;
        const bool is_varying = !fold || (!(st0.flags & 1) && st0.src_col >= 0);
        const int p = st0.parent, c = st0.child;
        if (!is_varying) {
            const HXf M = host_constant_joint(st0);
            rel[c] = rel_identity[p] ? M : hxf_mul(rel[p], M);
            rel_identity[c] = 0;
            anchor[c] = anchor[p];
            if (pl.emit[c] >= 0) {
                ccdev::FkStep e;
                memset(&e, 0, sizeof(e));
                e.type = 1; e.parent = anchor[c]; e.child = c; e.src_col = -1; e.flags = 1 | 16; e.load_parent = -2; e.save = -1;
                hxf_store(rel[c], e.L);
                e.R[0] = 1.0;
                consts[anchor[c]].push_back(e);
            }
        } else {
            ccdev::FkStep v = st0;
            v.parent = fold ? anchor[p] : p;
            if (fold && !rel_identity[p]) { const HXf L2 = hxf_mul(rel[p], hxf_load(st0.L)); hxf_store(L2, v.L); }
            varying.push_back(v);
            anchor[c] = c; rel_identity[c] = 1;
        }
    }
    for (const ccdev::FkStep& e : consts[0]) pl.steps.push_back(e);          // rigidly attached to the base link
    for (const ccdev::FkStep& v : varying) {
        pl.steps.push_back(v);
        for (const ccdev::FkStep& e : consts[v.child]) pl.steps.push_back(e);
    }
    // two-phase kernel: every varying step's child pose has a slot in the per-configuration pose table (0 = link 0)
    {
        std::vector<int> pose_of_link(nl, 0);
        int nv = 0;
        for (ccdev::FkStep& st : pl.steps) {
            st.pose_in = pose_of_link[st.parent];
            if (st.flags & 16) st.pose_out = -1;
            else { st.pose_out = ++nv; pose_of_link[st.child] = nv; }
        }
        pl.npose = 1 + nv;
    }
    // where each varying step finds its starting pose: the running transform, the base, or a parked pose
    std::vector<int> save_of_link(nl, -1);
    int last_varying = -1;
    for (size_t i = 0; i < pl.steps.size(); ++i) {
        ccdev::FkStep& st = pl.steps[i];
        if (st.flags & 16) continue;
        st.save = -1;
        if (last_varying >= 0 && pl.steps[last_varying].child == st.parent) st.load_parent = -2;
        else if (st.parent == 0) st.load_parent = -1;
        else {
            if (save_of_link[st.parent] < 0) {
                save_of_link[st.parent] = pl.nsave++;
                for (size_t j = 0; j < i; ++j) if (!(pl.steps[j].flags & 16) && pl.steps[j].child == st.parent) pl.steps[j].save = save_of_link[st.parent];
            }
            st.load_parent = save_of_link[st.parent];
        }
        last_varying = (int)i;
    }
}

static int build_chain_steps(const std::vector<cc_joint>& joints, int nl, const std::vector<int>& col_of_dof, int nactive, const double* dofvals,
                             std::vector<ccdev::FkStep>& steps, std::vector<int>& depth);

// One configuration through a compiled chain program on the host, in the order the device walks it.  walk 0: cc_fk_kernel
// (ccdev::fk_chain: running transform, poses later branches start from parked in `save` slots); walk 1: cc_fk2_kernel (every varying
// step reads and writes a slot of the per-configuration pose table, the constant steps hang off those slots).
static void host_run_program(const Plan& pl, const double* base8, const double* qrow, int nl, int walk, double* out)
{
    for (size_t i = 0; i < (size_t)nl * 8; ++i) out[i] = 0.0;
    const HXf base = hxf_load(base8);
    hxf_store(base, out);
    if (walk == 0) {
        std::vector<HXf> saved((size_t)std::max(1, pl.nsave));
        HXf cur = base;
        for (const ccdev::FkStep& st : pl.steps) {
            if (st.load_parent == -1) cur = base;
            else if (st.load_parent >= 0) cur = saved[st.load_parent];
            if (st.flags & 16) { if (pl.emit[st.child] >= 0) hxf_store(hxf_mul(cur, hxf_load(st.L)), out + 8 * (size_t)st.child); continue; }
            cur = hxf_mul(cur, host_joint(st, qrow));
            if (st.save >= 0) saved[st.save] = cur;
            if (pl.emit[st.child] >= 0) hxf_store(cur, out + 8 * (size_t)st.child);
        }
    } else {
        std::vector<HXf> pose((size_t)std::max(1, pl.npose));
        pose[0] = base;
        for (const ccdev::FkStep& st : pl.steps) {
            if (st.flags & 16) continue;
            pose[st.pose_out] = hxf_mul(pose[st.pose_in], host_joint(st, qrow));
            if (pl.emit[st.child] >= 0) hxf_store(pose[st.pose_out], out + 8 * (size_t)st.child);
        }
        for (const ccdev::FkStep& st : pl.steps)
            if ((st.flags & 16) && pl.emit[st.child] >= 0) hxf_store(hxf_mul(pose[st.pose_in], hxf_load(st.L)), out + 8 * (size_t)st.child);
    }
}

int cc_debug_fk_program(const cc_joint* joints, int njoints, int nlinks, int ndof, const double* dofvals, const double* base_pose,
                        const int32_t* active, int nactive, const uint8_t* wanted_links, int walk, const double* q, int64_t n,
                        double* out_poses, int32_t* out_counts)
{
    if (!joints || njoints <= 0 || nlinks <= 0 || ndof < 0 || (ndof && !dofvals) || !base_pose || nactive < 0 || (nactive && !active) ||
        (walk != 0 && walk != 1) || n < 0 || (n && (!q || !out_poses)))
        return fail(CC_ERR_INVALID, "cc_debug_fk_program: bad arguments");
    std::vector<cc_joint> J(joints, joints + njoints);
    std::vector<int> col_of_dof((size_t)ndof, -1);
    for (int c = 0; c < nactive; ++c) {
        if (active[c] < 0 || active[c] >= ndof) return fail(CC_ERR_INVALID, "cc_debug_fk_program: active DOF index out of range");
        col_of_dof[active[c]] = c;
    }
    int nexternal = 0;
    for (const cc_joint& j : J) {
        if (j.parent_link < 0 || j.parent_link >= nlinks || j.child_link < 0 || j.child_link >= nlinks) return fail(CC_ERR_INVALID, "joint link out of range");
        if (j.dof_index >= ndof || j.mimic_dof >= ndof) return fail(CC_ERR_INVALID, "joint dof out of range");
        if (j.mimic_dof == CC_MIMIC_EXTERNAL) nexternal = std::max(nexternal, j.reserved + 1);
    }
    std::unique_ptr<cc_query> Q(new cc_query());
    Q->nq = nlinks;
    Q->has_fk = true;
    std::vector<int> depth;
    const int rc = build_chain_steps(J, nlinks, col_of_dof, nactive, dofvals, Q->steps, depth);
    if (rc != CC_OK) return rc;
    Plan pl;
    pl.emit.assign((size_t)nlinks, -1);
    int ne = 0;
    for (int l = 0; l < nlinks; ++l) if (!wanted_links || wanted_links[l]) pl.emit[l] = ne++;
    compile_fk_program(Q.get(), pl);
    if (out_counts) {
        int nv = 0;
        for (const ccdev::FkStep& st : pl.steps) nv += (st.flags & 16) ? 0 : 1;
        out_counts[0] = (int32_t)pl.steps.size(); out_counts[1] = nv; out_counts[2] = pl.npose; out_counts[3] = pl.nsave;
    }
    const int ncols = nactive + nexternal;
    for (int64_t i = 0; i < n; ++i)
        host_run_program(pl, base_pose, q + (size_t)i * ncols, nlinks, walk, out_poses + (size_t)i * nlinks * 8);
    return CC_OK;
}

// Lay one pair list out for the device.  `dynamic[l]`: query link l gets a per-configuration pose (scratch slot); every
// other link a pair names (environment links, query links no active DOF moves) becomes a static slot at its current pose.
static int build_plan(cc_query* Q, Plan& pl, const std::vector<int>& plist, const std::vector<uint8_t>& dynamic, bool with_fk)
{
    cc_scene* S = Q->scene;
    HBody* B = S->bodies[Q->body];
    const int nl = Q->nq;
    pl.emit.assign(std::max(1, nl), -1);
    pl.ndyn = 0;
    std::vector<uint8_t> dyn(dynamic);
    if (&pl == &Q->mainl && with_fk && Q->has_fk)      // the links of the analytic pairs need their per-configuration poses too
        for (int p : Q->p_analytic) {
            const int a = Q->quads[4 * p];
            if (Q->moves[a]) dyn[a] = 1;
            if (Q->quads[4 * p + 3] == 1) { const int b2 = Q->quads[4 * p + 2]; if (Q->moves[b2]) dyn[b2] = 1; }
        }
    for (int l = 0; l < nl; ++l) if (dyn[l]) pl.emit[l] = pl.ndyn++;
    std::vector<std::pair<int, int>> statics;                // (body, link) in order of first use
    auto static_slot = [&](int body, int link) {
        for (size_t i = 0; i < statics.size(); ++i) if (statics[i].first == body && statics[i].second == link) return (int)i;
        statics.push_back(std::make_pair(body, link));
        return (int)statics.size() - 1;
    };
    std::vector<std::pair<int, int>> enc;                    // (slot_a, static index or -1 - dyn index) resolved after counting
    pl.pairs.clear(); pl.orig.clear(); pl.nenv = 0;
    std::vector<int> sa_v, sb_v;
    for (int p : plist) {
        const int a = Q->quads[4 * p], ob = Q->quads[4 * p + 1], j = Q->quads[4 * p + 2], cls = Q->quads[4 * p + 3];
        const int sa = pl.emit[a] >= 0 ? pl.emit[a] : -1 - static_slot(Q->ql_body[a], Q->ql_link[a]);
        int sb;
        if (cls == 1) sb = pl.emit[j] >= 0 ? pl.emit[j] : -1 - static_slot(Q->ql_body[j], Q->ql_link[j]);
        else sb = -1 - static_slot(ob, j);
        sa_v.push_back(sa); sb_v.push_back(sb);
        pl.orig.push_back(p);
        if (cls == 0) pl.nenv++;
    }
    // analytic pairs (cc_analytic_kernel) live in the same slot space
    const bool with_analytic = (&pl == &Q->mainl || &pl == &Q->full) && !Q->p_analytic.empty();
    std::vector<int> asa_v, asb_v;
    pl.apairs.clear(); pl.aorig.clear(); pl.anenv = 0;
    if (with_analytic)
        for (int p : Q->p_analytic) {
            const int a = Q->quads[4 * p], ob = Q->quads[4 * p + 1], j = Q->quads[4 * p + 2], cls = Q->quads[4 * p + 3];
            const int sa = pl.emit[a] >= 0 ? pl.emit[a] : -1 - static_slot(Q->ql_body[a], Q->ql_link[a]);
            int sb;
            if (cls == 1) sb = pl.emit[j] >= 0 ? pl.emit[j] : -1 - static_slot(Q->ql_body[j], Q->ql_link[j]);
            else sb = -1 - static_slot(ob, j);
            asa_v.push_back(sa); asb_v.push_back(sb);
            pl.aorig.push_back(p);
            if (cls == 0) pl.anenv++;
        }
    pl.nstatic = (int)statics.size();
    pl.statics = statics;
    if (pl.ndyn + pl.nstatic > 0xffff) return fail(CC_ERR_UNSUPPORTED, "more than 65535 link slots");
    for (size_t i = 0; i < sa_v.size(); ++i) {
        const uint32_t sa = sa_v[i] >= 0 ? (uint32_t)sa_v[i] : (uint32_t)(pl.ndyn + (-1 - sa_v[i]));
        const uint32_t sb = sb_v[i] >= 0 ? (uint32_t)sb_v[i] : (uint32_t)(pl.ndyn + (-1 - sb_v[i]));
        pl.pairs.push_back(sa | (sb << 16));
    }
    for (size_t i = 0; i < asa_v.size(); ++i) {
        const uint32_t sa = asa_v[i] >= 0 ? (uint32_t)asa_v[i] : (uint32_t)(pl.ndyn + (-1 - asa_v[i]));
        const uint32_t sb = asb_v[i] >= 0 ? (uint32_t)asb_v[i] : (uint32_t)(pl.ndyn + (-1 - asb_v[i]));
        pl.apairs.push_back(sa | (sb << 16));
    }
    std::vector<int32_t> roots((size_t)std::max(1, pl.ndyn + pl.nstatic), 0);
    std::vector<float> linkbox((size_t)std::max(1, pl.ndyn) * 8, 0.f), stxf((size_t)std::max(1, pl.nstatic) * 12, 0.f),
        staabb((size_t)std::max(1, pl.nstatic) * 8, 0.f);
    for (int l = 0; l < nl; ++l) {
        if (pl.emit[l] < 0) continue;
        const HLink& L = Q->qlink(l);
        roots[pl.emit[l]] = L.node_base < 0 ? 0 : L.node_base;
        float* o = &linkbox[(size_t)pl.emit[l] * 8];
        if (!L.nodes.empty()) { for (int k = 0; k < 3; ++k) { o[k] = L.nodes[0].c[k]; o[3 + k] = L.nodes[0].e[k]; } o[6] = 1.0f; }
        else { o[0] = o[1] = o[2] = 0.f; o[3] = o[4] = o[5] = -1e30f; }   // no geometry: an AABB that overlaps nothing
    }
    pl.st_pose.assign((size_t)pl.nstatic * 8, 0.0);
    for (int i = 0; i < pl.nstatic; ++i) {
        const HLink& L = S->bodies[statics[i].first]->links[statics[i].second];
        roots[pl.ndyn + i] = L.node_base < 0 ? 0 : L.node_base;
        pose_to_rt32(L.xf, Q->center, &stxf[(size_t)i * 12]);
        static_aabb(L, Q->center, &staabb[(size_t)i * 8]);
        memcpy(&pl.st_pose[(size_t)i * 8], L.xf, 8 * sizeof(double));
    }
    if (with_analytic) {
        // primitives of every slot (dynamic query links first, then the statics), FP32, pose in the link frame
        std::vector<int32_t> slot_prim(1, 0);
        std::vector<ccdev::Prim> prims;
        auto add_link = [&](const HLink& L) {
            for (const cc_geom& g : L.prims) {
                ccdev::Prim d;
                double Rm[9];
                quat_to_matrix(g.pose, Rm);
                for (int k = 0; k < 9; ++k) d.R[k] = (float)Rm[k];
                for (int k = 0; k < 3; ++k) d.t[k] = (float)g.pose[4 + k];
                d.dims[0] = (float)g.dims[0]; d.dims[1] = (float)g.dims[1]; d.dims[2] = (float)g.dims[2];
                if (g.type == CC_GEOM_CYLINDER || g.type == CC_GEOM_CAPSULE) d.dims[1] = (float)(0.5 * g.dims[1]);     // half height
                d.type = g.type;
                prims.push_back(d);
            }
            slot_prim.push_back((int32_t)prims.size());
        };
        std::vector<int> link_of_slot(pl.ndyn, -1);
        for (int l = 0; l < nl; ++l) if (pl.emit[l] >= 0) link_of_slot[pl.emit[l]] = l;
        for (int d = 0; d < pl.ndyn; ++d) add_link(Q->qlink(link_of_slot[d]));
        for (int i = 0; i < pl.nstatic; ++i) add_link(S->bodies[statics[i].first]->links[statics[i].second]);
        CC_CUDA(pl.d_apairs.upload(pl.apairs)); CC_CUDA(pl.d_aorig.upload(pl.aorig));
        CC_CUDA(pl.d_slot_prim.upload(slot_prim)); CC_CUDA(pl.d_prims.upload(prims));
    }
    pl.steps.clear();
    pl.nsave = 0;
    if (with_fk) compile_fk_program(Q, pl);
    CC_CUDA(pl.d_pairs.upload(pl.pairs)); CC_CUDA(pl.d_orig.upload(pl.orig)); CC_CUDA(pl.d_emit.upload(pl.emit));
    CC_CUDA(pl.d_slot_root.upload(roots)); CC_CUDA(pl.d_link_box.upload(linkbox));
    CC_CUDA(pl.d_st_xf.upload(stxf)); CC_CUDA(pl.d_st_aabb.upload(staabb));
    CC_CUDA(pl.d_steps.upload(pl.steps));
    // split tables: part k of S takes every S-th block of 32 pairs of each class (env first), reference order kept
    for (int si = 0; si < 3; ++si) {
        const int S = 2 << si;
        std::vector<uint32_t> sp; std::vector<int32_t> so, pb(1, 0), pn;
        const int np = (int)pl.pairs.size();
        for (int k = 0; k < S; ++k) {
            int ne = 0;
            for (int i = 0; i < pl.nenv; ++i) if ((i / 32) % S == k) { sp.push_back(pl.pairs[i]); so.push_back(pl.orig[i]); ++ne; }
            for (int i = pl.nenv; i < np; ++i) if (((i - pl.nenv) / 32) % S == k) { sp.push_back(pl.pairs[i]); so.push_back(pl.orig[i]); }
            pn.push_back(ne);
            pb.push_back((int32_t)sp.size());
        }
        pl.split[si].orig = so;
        CC_CUDA(pl.split[si].d_pairs.upload(sp)); CC_CUDA(pl.split[si].d_orig.upload(so));
        CC_CUDA(pl.split[si].d_part_begin.upload(pb)); CC_CUDA(pl.split[si].d_part_nenv.upload(pn));
    }
    return CC_OK;
}

// Evaluate the pose-invariant pairs once, with the traversal kernel, at the current DOF values; the result is merged into
// every batched verdict as a constant mask.
static int evaluate_hoisted(cc_query* Q)
{
    cc_scene* S = Q->scene;
    const int np = (int)Q->quads.size() / 4;
    int rc = CC_OK;
    Q->hoisted_first_env = 0x7fffffff; Q->hoisted_first_self = 0x7fffffff;
    Q->hoisted_mask.assign(Q->mask_words, 0u);
    CC_CUDA(Q->d_hoisted_mask.upload(Q->hoisted_mask));
    if (!Q->hoisted.pairs.empty()) {
        CC_CUDA(Q->d_q.reserve(std::max<size_t>(1, Q->cur_active.size())));
        if (!Q->cur_active.empty())
            CC_CUDA(cudaMemcpy(Q->d_q.p, Q->cur_active.data(), Q->cur_active.size() * sizeof(double), cudaMemcpyHostToDevice));
        CC_CUDA(Q->d_mask.reserve(Q->mask_words));
        rc = launch(Q, Q->hoisted, false, Q->d_q.p, 1, CC_BATCH_ENV | CC_BATCH_SELF | CC_BATCH_ALL_PAIRS, nullptr, nullptr,
                        nullptr, Q->d_mask.p, nullptr, 0);
        if (rc != CC_OK) return rc;
        CC_CUDA(cudaMemcpy(Q->hoisted_mask.data(), Q->d_mask.p, Q->mask_words * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        rc = check_err_flag(S);
        if (rc != CC_OK) return rc;
        S->stats.configs -= 1;
        for (int p = 0; p < np; ++p)
            if (Q->hoisted_mask[p >> 5] & (1u << (p & 31))) {
                if (p < Q->nenv_total) { if (p < Q->hoisted_first_env) Q->hoisted_first_env = p; }
                else if (p < Q->hoisted_first_self) Q->hoisted_first_self = p;
            }
        CC_CUDA(Q->d_hoisted_mask.upload(Q->hoisted_mask));
        CC_CUDA(cudaDeviceSynchronize());
    }
    return CC_OK;
}

// The chain program of a body over one active set: one ccdev::FkStep per joint, depth-first (children in listed order) so that a
// child usually follows its parent.  `depth[l]` comes back -1 for a link no joint reaches.  (Shared by the query builder and the
// device-free cc_debug_fk_program.)
static int build_chain_steps(const std::vector<cc_joint>& joints, int nl, const std::vector<int>& col_of_dof, int nactive, const double* dofvals,
                             std::vector<ccdev::FkStep>& steps, std::vector<int>& depth)
{
    steps.clear();
    depth.assign((size_t)nl, -1);
    depth[0] = 0;
    std::vector<int> sdepth(joints.size(), 0);
    for (size_t k = 0; k < joints.size(); ++k) {
        const cc_joint& j = joints[k];
        if (depth[j.parent_link] < 0) return fail(CC_ERR_INVALID, "joints are not listed parent-before-child from link 0");
        if (depth[j.child_link] >= 0) return fail(CC_ERR_UNSUPPORTED, "closed kinematic loops are not supported");
        depth[j.child_link] = depth[j.parent_link] + 1;
        sdepth[k] = depth[j.child_link];
    }
    // depth-first order (children in listed order): a child usually follows its parent, so the device keeps the
    // running transform in registers and parks only the poses later branches start from
    std::vector<std::vector<int>> kids(nl);
    for (size_t k = 0; k < joints.size(); ++k) kids[joints[k].parent_link].push_back((int)k);
    std::vector<int> ord;
    {
        std::vector<std::pair<int, size_t>> stk;     // (link, next child position)
        stk.push_back(std::make_pair(0, (size_t)0));
        while (!stk.empty()) {
            std::pair<int, size_t>& top = stk.back();
            if (top.second >= kids[top.first].size()) { stk.pop_back(); continue; }
            const int k = kids[top.first][top.second++];
            ord.push_back(k);
            stk.push_back(std::make_pair(joints[k].child_link, (size_t)0));
        }
    }
    if (ord.size() != joints.size()) return fail(CC_ERR_INVALID, "joint tree is not connected to link 0");
    for (int k : ord) {
        const cc_joint& j = joints[k];
        ccdev::FkStep st;
        memset(&st, 0, sizeof(st));
        st.type = j.type; st.parent = j.parent_link; st.child = j.child_link;
        int srcdof = j.mimic_dof >= 0 ? j.mimic_dof : j.dof_index;
        st.src_col = srcdof >= 0 ? col_of_dof[srcdof] : -1;
        st.src_const = srcdof >= 0 ? dofvals[srcdof] : 0.0;
        st.flags = (j.is_static ? 1 : 0) | (j.mimic_dof >= 0 ? 2 : 0) | (j.is_circular ? 4 : 0);
        st.src_lower = j.lower; st.src_upper = j.upper;
        bool src_circ = j.is_circular != 0;
        if (j.mimic_dof >= 0) {
            // limits of the DOF the equation reads
            bool found = false;
            for (const cc_joint& o : joints)
                if (o.dof_index == j.mimic_dof) { st.src_lower = o.lower; st.src_upper = o.upper; src_circ = o.is_circular != 0; found = true; break; }
            if (!found) return fail(CC_ERR_INVALID, "mimic source DOF has no joint");
        } else if (srcdof < 0) {
            // passive, non-mimic joint: value 0 clamped to its own limits (kinbody.cpp:2262-2275)
            st.src_const = 0.0;
        }
        if (src_circ) st.flags |= 8;
        st.a = j.mimic_a; st.b = j.mimic_b; st.lower = j.lower; st.upper = j.upper;
        if (j.mimic_dof == CC_MIMIC_EXTERNAL) {
            // the equation is evaluated by the caller (fparser expressions the reference evaluates on the host too,
            // kinbody.cpp:2308-2320): its value arrives as a column of q and goes through the mimic limit rules (:2326-2386)
            st.src_col = nactive + j.reserved;
            st.src_const = 0.0;
            st.flags |= 2 | 8;                       // mimic; the "source" is the value itself: no source clamp
            st.a = 1.0; st.b = 0.0;
        }
        memcpy(st.L, j.left, sizeof(st.L)); memcpy(st.R, j.right, sizeof(st.R)); memcpy(st.axis, j.axis, sizeof(st.axis));
        steps.push_back(st);
    }
    return CC_OK;
}

// What a batch can change.  `moves[l]`: link l's WORLD pose depends on the rows of q — an active DOF affects it
// (KinBody::DoesAffect, which follows mimic equations to their source DOFs), the base itself moves (affine DOFs), or an
// externally evaluated mimic joint (CC_MIMIC_EXTERNAL: a new value with every row whatever the active set is) sits above it.
// `pose_invariant(a, b)`: the pose of link b in link a's frame cannot change — no joint on the tree path between them is driven
// by an active DOF, directly, through a mimic equation, or by the caller.  (Comparing the DOFs that affect the two links is not
// enough: a mimic joint between them may follow a DOF that moves both.)
struct MotionClasses {
    std::vector<uint8_t> moves;          // per link
    std::vector<int> parent_joint;       // per link: the joint whose child it is, -1 for the root
    std::vector<uint8_t> jdriven;        // per joint
    const std::vector<cc_joint>* joints = nullptr;

    void build(const std::vector<cc_joint>& J, int nl, const std::vector<int>& act, const std::vector<int>& col_of_dof, bool base_moves,
               const std::function<bool(int, int)>& affects)
    {
        joints = &J;
        moves.assign((size_t)nl, 0);
        for (int l = 0; l < nl; ++l)
            for (int d : act) if (affects(d, l)) { moves[l] = 1; break; }
        if (base_moves) std::fill(moves.begin(), moves.end(), (uint8_t)1);
        parent_joint.assign((size_t)nl, -1);
        jdriven.assign(J.size(), 0);
        // (joints are listed parent before child; the chain compiler rejects any other order)
        for (size_t k = 0; k < J.size(); ++k) {
            const cc_joint& j = J[k];
            if (j.child_link < 0 || j.child_link >= nl || j.parent_link < 0 || j.parent_link >= nl) continue;
            parent_joint[j.child_link] = (int)k;
            if (j.mimic_dof == CC_MIMIC_EXTERNAL || moves[j.parent_link]) moves[j.child_link] = 1;
            if (j.is_static) continue;
            const int srcdof = j.mimic_dof >= 0 ? j.mimic_dof : j.dof_index;
            jdriven[k] = j.mimic_dof == CC_MIMIC_EXTERNAL || (srcdof >= 0 && srcdof < (int)col_of_dof.size() && col_of_dof[srcdof] >= 0);
        }
    }
    bool pose_invariant(int a, int b) const
    {
        const std::vector<cc_joint>& J = *joints;
        std::vector<int> pa, pb;
        for (int l = a, g = 0; l >= 0 && parent_joint[l] >= 0 && g < 1 << 16; ++g) { pa.push_back(parent_joint[l]); l = J[parent_joint[l]].parent_link; }
        for (int l = b, g = 0; l >= 0 && parent_joint[l] >= 0 && g < 1 << 16; ++g) { pb.push_back(parent_joint[l]); l = J[parent_joint[l]].parent_link; }
        while (!pa.empty() && !pb.empty() && pa.back() == pb.back()) { pa.pop_back(); pb.pop_back(); }
        for (int k : pa) if (jdriven[k]) return false;
        for (int k : pb) if (jdriven[k]) return false;
        return true;
    }
};

static int finish_query(cc_query* Q, const int32_t* active, int nactive, bool all_dofs_active, bool allow_hoist)
{
    cc_scene* S = Q->scene;
    HBody* B = S->bodies[Q->body];
    const int nl = Q->nq;
    const int nlb = (int)B->links.size();
    // the joint list of the query: the body's joints, then one static joint per grabbed link hanging off its grabbing link
    // (KinBody::_UpdateGrabbedBodies: T(grabbed link) = T(grabbing link) * relative pose)
    std::vector<cc_joint> joints(B->joints);
    for (int l = nlb; l < nl; ++l) {
        const HBody::Grab& g = B->grabs[Q->ql_grab[l]];
        cc_joint j;
        memset(&j, 0, sizeof(j));
        j.type = CC_JOINT_REVOLUTE; j.parent_link = g.link; j.child_link = l; j.dof_index = -1; j.mimic_dof = -1; j.is_static = 1;
        memcpy(j.left, &g.rel[8 * (size_t)Q->ql_link[l]], 8 * sizeof(double));
        j.right[0] = 1.0;
        joints.push_back(j);
    }
    // ---- which links move with the active DOFs
    std::vector<int> col_of_dof(B->ndof, -1);
    std::vector<int> act;
    if (all_dofs_active) { for (int d = 0; d < B->ndof; ++d) act.push_back(d); }
    else act.assign(active, active + nactive);
    for (size_t c = 0; c < act.size(); ++c) {
        if (act[c] < 0 || act[c] >= B->ndof) return fail(CC_ERR_INVALID, "active DOF index out of range");
        col_of_dof[act[c]] = (int)c;
    }
    // columns of q: active joint DOFs | values of externally evaluated mimic joints (CC_MIMIC_EXTERNAL, by ordinal) | affine DOFs
    Q->nexternal = 0;
    for (const cc_joint& j : B->joints) if (j.mimic_dof == CC_MIMIC_EXTERNAL) Q->nexternal = std::max(Q->nexternal, j.reserved + 1);
    Q->ncols = (int)act.size() + Q->nexternal + Q->naffine;
    Q->col_of_dof = col_of_dof;
    Q->cur_active.assign((size_t)Q->ncols, 0.0);
    for (size_t c = 0; c < act.size(); ++c) Q->cur_active[c] = B->dofvals[act[c]];
    if (Q->naffine) {
        // the bound pose in affine coordinates — only the translation matters: the pairs evaluated at the bound values are the
        // pose-invariant (self) ones, and a base rotation does not change them
        size_t c = act.size() + (size_t)Q->nexternal;
        for (int k = 0; k < 3; ++k) if (Q->affine_mask & (1 << k)) Q->cur_active[c++] = B->links.empty() ? 0.0 : B->links[0].xf[4 + k];
        if (Q->affine_mask & CC_DOF_RotationQuat) Q->cur_active[c] = 1.0;
    }
    auto affects = [&](int d, int l) { return B->affects[(size_t)d * nlb + Q->ql_host[l]] != 0; };
    MotionClasses motion;
    motion.build(joints, nl, act, col_of_dof, Q->naffine != 0, affects);
    const std::vector<uint8_t>& moves = motion.moves;

    // ---- FK chain program
    Q->has_fk = true;
    Q->steps.clear();
    if (!joints.empty()) {
        std::vector<int> depth;
        const int rcs = build_chain_steps(joints, nl, col_of_dof, (int)act.size(), B->dofvals.data(), Q->steps, depth);
        if (rcs != CC_OK) return rcs;
        for (int l = 1; l < nl; ++l)
            if (depth[l] < 0 && Q->qlink(l).ntris() > 0) Q->has_fk = false;   // a link no joint reaches: pose mode only
    } else {
        for (int l = 1; l < nl; ++l)
            if (Q->qlink(l).ntris() > 0) Q->has_fk = false;
    }
    memcpy(Q->base, B->links.empty() ? std::vector<double>(8, 0).data() : B->links[0].xf, sizeof(Q->base));
    Q->center[0] = Q->base[4]; Q->center[1] = Q->base[5]; Q->center[2] = Q->base[6];

    // ---- split the pair list and lay the three plans out
    const int np = (int)Q->quads.size() / 4;
    Q->mask_words = std::max(1, (np + 31) / 32);
    std::vector<int> p_full, p_full_mesh, p_main, p_hoist;
    Q->p_analytic.clear();
    for (int p = 0; p < np; ++p) {
        if (!Q->pair_live[p]) continue;
        const int a = Q->quads[4 * p], cls = Q->quads[4 * p + 3];
        p_full.push_back(p);
        if (S->analytic) {
            // fclrave semantics: a pair with a link made of primitives is decided on the exact shapes (cc_analytic.cuh), per
            // configuration (never hoisted, never certified)
            const HLink& La = Q->qlink(a);
            const HLink& Lb = cls == 1 ? Q->qlink(Q->quads[4 * p + 2]) : S->bodies[Q->quads[4 * p + 1]]->links[Q->quads[4 * p + 2]];
            if (!La.prims.empty() || !Lb.prims.empty()) { Q->p_analytic.push_back(p); continue; }
        }
        p_full_mesh.push_back(p);
        bool hoist = false;
        if (allow_hoist && Q->has_fk) {
            if (cls == 0) hoist = !moves[a];
            else {
                hoist = motion.pose_invariant(a, Q->quads[4 * p + 2]);
            }
        }
        (hoist ? p_hoist : p_main).push_back(p);
    }
    // minimum-distance queries: pairs between links that no active DOF moves keep their distance for the whole batch
    Q->p_dist_main.clear(); Q->p_dist_static.clear(); Q->dist_dynamic.assign(nl, 1);
    Q->dist_main_ready = false;
    for (auto& sm : Q->dist_static_min) sm.valid = false;
    {
        const bool split_static = allow_hoist && Q->has_fk && !Q->naffine;
        if (split_static) Q->dist_dynamic.assign(nl, 0);
        for (int p : p_full) {
            const int a = Q->quads[4 * p], cls = Q->quads[4 * p + 3];
            const int b = cls == 1 ? Q->quads[4 * p + 2] : -1;
            if (split_static && !moves[a] && (b < 0 || !moves[b])) { Q->p_dist_static.push_back(p); continue; }
            Q->p_dist_main.push_back(p);
            if (split_static) { if (moves[a]) Q->dist_dynamic[a] = 1; if (b >= 0 && moves[b]) Q->dist_dynamic[b] = 1; }
        }
    }
    std::vector<uint8_t> all_dynamic(nl, 1), main_dynamic(nl, 0);
    for (int p : p_main) {
        const int a = Q->quads[4 * p];
        if (moves[a]) main_dynamic[a] = 1;
        if (Q->quads[4 * p + 3] == 1) { const int b = Q->quads[4 * p + 2]; if (moves[b]) main_dynamic[b] = 1; }
    }
    CC_CUDA(cudaSetDevice(S->device));
    Q->p_main = p_main;
    Q->p_main_all = p_main;
    Q->p_hoist = p_hoist;
    Q->moves = moves;
    Q->p_full = p_full;
    Q->allow_tables = allow_hoist && Q->has_fk && Q->nq == (int)B->links.size() && !Q->naffine;     // not with grabbed links (their chain is not walked below), not with a moving base
    Q->pose_epoch = S->pose_epoch;           // everything below reads the live scene state: that IS the query's snapshot
    int rc = build_plan(Q, Q->full, p_full_mesh, all_dynamic, false);
    if (rc == CC_OK) rc = build_plan(Q, Q->hoisted, p_hoist, all_dynamic, true);
    if (rc == CC_OK) rc = build_plan(Q, Q->mainl, p_main, Q->has_fk ? main_dynamic : all_dynamic, true);
    if (rc != CC_OK) return rc;
    rc = evaluate_hoisted(Q);
    if (rc != CC_OK) return rc;
    return CC_OK;
}

struct AffineSpec { int mask = 0; double axis[3] = {0, 0, 1}; double qstart[4] = {1, 0, 0, 0}; };
static int query_create_one(cc_scene* s, int body, const int32_t* active_dofs, int nactive, int options, cc_query** out, const AffineSpec* affine = nullptr);
static int query_create_pairs_one(cc_scene* s, int body, const int32_t* triples, int ntriples, cc_query** out);

}  // extern "C"
// bind the same query on every peer replica of a multi-device scene
template <typename F>
static int bind_on_peers(cc_scene* s, cc_query* Q, F create)
{
    for (cc_scene* p : s->peers) {
        cc_query* pq = nullptr;
        const int rc = create(p, &pq);
        if (rc != CC_OK) { const std::string keep = g_err; cc_query_destroy(Q); g_err = keep; return rc; }
        Q->peers.push_back(pq);
    }
    if (!s->peers.empty()) CC_CUDA(cudaSetDevice(s->device));
    return CC_OK;
}
extern "C" {

int cc_query_create(cc_scene* s, int body, const int32_t* active_dofs, int nactive, int options, cc_query** out)
{
    if (!s || !out) return fail(CC_ERR_INVALID, "bad arguments");
    cc_query* Q = nullptr;
    int rc = query_create_one(s, body, active_dofs, nactive, options, &Q);
    if (rc != CC_OK) return rc;
    rc = bind_on_peers(s, Q, [&](cc_scene* p, cc_query** o) { return query_create_one(p, body, active_dofs, nactive, options, o); });
    if (rc != CC_OK) return rc;
    *out = Q;
    return CC_OK;
}

int cc_query_create_pairs(cc_scene* s, int body, const int32_t* triples, int ntriples, cc_query** out)
{
    if (!s || !out) return fail(CC_ERR_INVALID, "bad arguments");
    cc_query* Q = nullptr;
    int rc = query_create_pairs_one(s, body, triples, ntriples, &Q);
    if (rc != CC_OK) return rc;
    // explicit pair lists serve the single-state overloads and the certificate builder: one configuration, one device
    *out = Q;
    return CC_OK;
}

int cc_query_create_affine(cc_scene* s, int body, const int32_t* active_dofs, int nactive, int affine_dofs, const double* rotation_axis3,
                           const double* quat_limit_start4, int options, cc_query** out)
{
    if (!s || !out || affine_dofs < 0 || affine_dofs > 63) return fail(CC_ERR_INVALID, "bad arguments");
    const int nrot = ((affine_dofs & CC_DOF_RotationAxis) ? 1 : 0) + ((affine_dofs & CC_DOF_Rotation3D) ? 1 : 0) + ((affine_dofs & CC_DOF_RotationQuat) ? 1 : 0);
    if (nrot > 1) return fail(CC_ERR_INVALID, "only one rotation parameterisation may be active");
    if ((affine_dofs & CC_DOF_RotationAxis) && !rotation_axis3) return fail(CC_ERR_INVALID, "DOF_RotationAxis needs its axis");
    if (affine_dofs == 0) return cc_query_create(s, body, active_dofs, nactive, options, out);
    if (active_dofs == nullptr && nactive == 0) {
        // (an empty joint list with affine DOFs is legitimate: pure base motion; NULL means "all joints" for cc_query_create only)
        static const int32_t none = 0;
        active_dofs = &none;
    }
    AffineSpec A;
    A.mask = affine_dofs;
    if (rotation_axis3) {
        const double len = sqrt(rotation_axis3[0] * rotation_axis3[0] + rotation_axis3[1] * rotation_axis3[1] + rotation_axis3[2] * rotation_axis3[2]);
        if (!(len > 0)) return fail(CC_ERR_INVALID, "zero rotation axis");
        for (int k = 0; k < 3; ++k) A.axis[k] = rotation_axis3[k];
    }
    if (quat_limit_start4) for (int k = 0; k < 4; ++k) A.qstart[k] = quat_limit_start4[k];
    cc_query* Q = nullptr;
    int rc = query_create_one(s, body, active_dofs, nactive, options, &Q, &A);
    if (rc != CC_OK) return rc;
    rc = bind_on_peers(s, Q, [&](cc_scene* p, cc_query** o) { return query_create_one(p, body, active_dofs, nactive, options, o, &A); });
    if (rc != CC_OK) return rc;
    *out = Q;
    return CC_OK;
}

static int query_create_one(cc_scene* s, int body, const int32_t* active_dofs, int nactive, int options, cc_query** out, const AffineSpec* affine)
{
    HBody* B = get_body(s, body);
    if (!B || !out || nactive < 0 || (nactive > 0 && !active_dofs)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!s->committed) return fail(CC_ERR_STATE, "cc_scene_commit has not been called since the last mesh change");
    const int nlb = (int)B->links.size();
    const bool all_active = active_dofs == nullptr;
    std::vector<int> act;
    if (all_active) for (int d = 0; d < B->ndof; ++d) act.push_back(d);
    else act.assign(active_dofs, active_dofs + nactive);
    for (int d : act) if (d < 0 || d >= B->ndof) return fail(CC_ERR_INVALID, "active DOF index out of range");
    cc_query* Q = new cc_query();
    Q->scene = s; Q->body = body;
    if (affine && affine->mask) {
        Q->affine_mask = affine->mask;
        Q->naffine = ((affine->mask & 1) ? 1 : 0) + ((affine->mask & 2) ? 1 : 0) + ((affine->mask & 4) ? 1 : 0) + ((affine->mask & CC_DOF_RotationAxis) ? 1 : 0) +
                     ((affine->mask & CC_DOF_Rotation3D) ? 3 : 0) + ((affine->mask & CC_DOF_RotationQuat) ? 4 : 0);
        for (int k = 0; k < 3; ++k) Q->affine_axis[k] = affine->axis[k];
        for (int k = 0; k < 4; ++k) Q->affine_qstart[k] = affine->qstart[k];
    }
    // query links: the body's own, then every grabbed body's in grab order
    for (int l = 0; l < nlb; ++l) { Q->ql_body.push_back(body); Q->ql_link.push_back(l); Q->ql_host.push_back(l); Q->ql_grab.push_back(-1); }
    std::vector<int> grab_base(B->grabs.size(), 0);
    for (size_t g = 0; g < B->grabs.size(); ++g) {
        HBody* G = get_body(s, B->grabs[g].body);
        if (!G || B->grabs[g].rel.size() != 8 * G->links.size()) { delete Q; return fail(CC_ERR_STATE, "grabbed body changed since cc_body_set_grabbed"); }
        grab_base[g] = (int)Q->ql_body.size();
        for (int l = 0; l < (int)G->links.size(); ++l) {
            Q->ql_body.push_back(B->grabs[g].body); Q->ql_link.push_back(l); Q->ql_host.push_back(B->grabs[g].link); Q->ql_grab.push_back((int)g);
        }
    }
    const int nl = (int)Q->ql_body.size();
    Q->nq = nl;
    if (nl > 0xffff) { delete Q; return fail(CC_ERR_UNSUPPORTED, "more than 65535 link slots"); }
    auto query_index = [&](int b, int l) -> int {       // (body, link) -> query link index, -1 if it is not one
        if (b == body) return l < nlb ? l : -1;
        for (size_t g = 0; g < B->grabs.size(); ++g)
            if (B->grabs[g].body == b) return l < (int)s->bodies[b]->links.size() ? grab_base[g] + l : -1;
        return -1;
    };
    // _IsActiveLink (collisionPQP.h:710-735): only links of the active robot itself are filtered, grabbed links never
    std::vector<uint8_t> active_link(nl, 1);
    const bool filter = (options & CC_CO_ActiveDOFs) && B->is_robot;
    if (filter && !(affine && affine->mask)) {           // with affine DOFs every link is active (collisionPQP.h:716-719)
        for (int l = 0; l < nlb; ++l) {
            active_link[l] = 0;
            for (int d : act) if (B->affects[(size_t)d * nlb + l]) { active_link[l] = 1; break; }
        }
    }
    auto in_attached_set = [&](int ob) {                 // KinBody::GetAttached is transitive: the body, its attachments, its grabs
        if (body_attached(*B, ob)) return true;
        for (const HBody::Grab& g : B->grabs) if (g.body == ob || body_attached(*s->bodies[g.body], ob)) return true;
        return false;
    };
    // env pairs: every attached body (the query body first, then its grabbed bodies in grab order; the reference walks a
    // pointer-ordered std::set, collisionPQP.h:314-329) x other bodies in environment order x its link x other link
    // (collisionPQP.h:448-487)
    for (int att = -1; att < (int)B->grabs.size(); ++att) {
        const int qb = att < 0 ? body : B->grabs[att].body;
        const int q0 = att < 0 ? 0 : grab_base[att];
        const HBody* QB = s->bodies[qb];
        for (int ob = 0; ob < (int)s->bodies.size(); ++ob) {
            HBody* O = s->bodies[ob];
            if (O->removed || in_attached_set(ob)) continue;
            for (int i = 0; i < (int)QB->links.size(); ++i) {
                if (!QB->links[i].enabled || !active_link[q0 + i] || QB->links[i].ntris() == 0) continue;
                for (int j = 0; j < (int)O->links.size(); ++j) {
                    if (!O->links[j].enabled || O->links[j].ntris() == 0) continue;
                    Q->quads.insert(Q->quads.end(), {q0 + i, ob, j, 0});
                    Q->pair_slot.push_back(0);
                    Q->pair_live.push_back(1);
                }
            }
        }
    }
    Q->nenv_total = (int)Q->quads.size() / 4;
    auto add_self = [&](int a, int b) {
        const HLink& La = Q->qlink(a); const HLink& Lb = Q->qlink(b);
        if (!La.enabled || !Lb.enabled || La.ntris() == 0 || Lb.ntris() == 0) return;
        if (!active_link[a] || !active_link[b]) return;              // DoPQP's _IsActiveLink on both links
        Q->quads.insert(Q->quads.end(), {a, body, b, 1});
        Q->pair_slot.push_back(b);
        Q->pair_live.push_back(1);
    };
    // self pairs: non-adjacent list, AO_Enabled [| AO_ActiveDOFs] (collisionPQP.h:364-386, robot.cpp:2023-2084)
    for (int32_t pk : B->nonadjacent) {
        int l0 = pk & 0xffff, l1 = (pk >> 16) & 0xffff;
        if (filter) {
            bool any = false;
            for (int d : act)
                if ((B->affects[(size_t)d * nlb + l0] != 0) != (B->affects[(size_t)d * nlb + l1] != 0)) { any = true; break; }
            if (!any) continue;
        }
        add_self(l0, l1);
    }
    // grabbed bodies (KinBody::CheckSelfCollision, kinbodycollision.cpp:83-205): per grabbed body, in grab order,
    //   its non-colliding links x its own links, its own non-adjacent pairs, then the later grabbed bodies link by link
    auto listed = [&](const HBody::Grab& g, int b, int l) {
        for (size_t k = 0; k + 1 < g.noncolliding.size(); k += 2) if (g.noncolliding[k] == b && g.noncolliding[k + 1] == l) return true;
        return false;
    };
    for (size_t g1 = 0; g1 < B->grabs.size(); ++g1) {
        const HBody::Grab& G1 = B->grabs[g1];
        const HBody* GB1 = s->bodies[G1.body];
        bool any_enabled = false;
        for (const HLink& L : GB1->links) any_enabled |= L.enabled;
        if (!any_enabled) continue;                                  // grabbedBody1.IsEnabled()
        for (size_t k = 0; k + 1 < G1.noncolliding.size(); k += 2) {
            const int a = query_index(G1.noncolliding[k], G1.noncolliding[k + 1]);
            if (a < 0) continue;                                     // a link of a body that is no longer grabbed
            for (int l = 0; l < (int)GB1->links.size(); ++l) add_self(a, grab_base[g1] + l);
        }
        for (int32_t pk : GB1->nonadjacent) add_self(grab_base[g1] + (pk & 0xffff), grab_base[g1] + ((pk >> 16) & 0xffff));
        for (size_t g2 = g1 + 1; g2 < B->grabs.size(); ++g2) {
            const HBody::Grab& G2 = B->grabs[g2];
            const HBody* GB2 = s->bodies[G2.body];
            bool en2 = false;
            for (const HLink& L : GB2->links) en2 |= L.enabled;
            if (!en2) continue;
            for (int l2 = 0; l2 < (int)GB2->links.size(); ++l2) {
                if (!GB2->links[l2].enabled || listed(G1, G2.body, l2)) continue;     // already checked through G1's list
                for (int l1 = 0; l1 < (int)GB1->links.size(); ++l1)
                    if (GB1->links[l1].enabled && listed(G2, G1.body, l1)) add_self(grab_base[g1] + l1, grab_base[g2] + l2);
            }
        }
    }
    Q->nself_total = (int)Q->quads.size() / 4 - Q->nenv_total;
    int rc = finish_query(Q, active_dofs, nactive, all_active, true);
    if (rc != CC_OK) { delete Q; return rc; }
    *out = Q;
    return CC_OK;
}

static int query_create_pairs_one(cc_scene* s, int body, const int32_t* triples, int ntriples, cc_query** out)
{
    HBody* B = get_body(s, body);
    if (!B || !out || ntriples < 0 || (ntriples > 0 && !triples)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!s->committed) return fail(CC_ERR_STATE, "cc_scene_commit has not been called since the last mesh change");
    const int nl = (int)B->links.size();
    cc_query* Q = new cc_query();
    Q->scene = s; Q->body = body; Q->nq = nl;
    std::vector<int> slot_body(nl, body), slot_link(nl);
    for (int l = 0; l < nl; ++l) slot_link[l] = l;
    for (int l = 0; l < nl; ++l) { Q->ql_body.push_back(body); Q->ql_link.push_back(l); Q->ql_host.push_back(l); Q->ql_grab.push_back(-1); }
    // env-class pairs first (index space [0, nenv)), then self-class pairs, each in the given order
    for (int pass = 0; pass < 2; ++pass) {
        for (int k = 0; k < ntriples; ++k) {
            int a = triples[3 * k], ob = triples[3 * k + 1], j = triples[3 * k + 2];
            HBody* O = get_body(s, ob);
            if (!O || a < 0 || a >= nl || j < 0 || j >= (int)O->links.size()) { delete Q; return fail(CC_ERR_INVALID, "bad pair triple"); }
            const bool self = ob == body;
            if ((pass == 0) == self) continue;
            Q->quads.insert(Q->quads.end(), {a, ob, j, self ? 1 : 0});
            // DoPQP: disabled links and links without a model never collide (collisionPQP.h:555-566);
            // such pairs keep their index but are not sent to the device
            bool live = B->links[a].enabled && O->links[j].enabled && B->links[a].ntris() > 0 && O->links[j].ntris() > 0;
            int slot = j;
            if (!self) {
                slot = -1;
                for (int sidx = nl; sidx < (int)slot_body.size(); ++sidx)
                    if (slot_body[sidx] == ob && slot_link[sidx] == j) { slot = sidx; break; }
                if (slot < 0) { slot = (int)slot_body.size(); slot_body.push_back(ob); slot_link.push_back(j); }
            }
            Q->pair_slot.push_back(slot);
            Q->pair_live.push_back(live ? 1 : 0);
        }
        if (pass == 0) Q->nenv_total = (int)Q->quads.size() / 4;
    }
    Q->nself_total = (int)Q->quads.size() / 4 - Q->nenv_total;
    if (slot_body.size() > 0xffff) { delete Q; return fail(CC_ERR_UNSUPPORTED, "more than 65535 link slots"); }
    int rc = finish_query(Q, nullptr, 0, true, false);
    if (rc != CC_OK) { delete Q; return rc; }
    *out = Q;
    return CC_OK;
}

// Forget free-space certificates of a query: those of the pairs flagged in `invalid` (all of them when it is NULL).  Pairs that
// had left the batched plan come back into it; the next certificate build (after another 16k configurations) examines only the
// pairs that have no certificate.
static int invalidate_certificates(cc_query* q, const std::vector<uint8_t>* invalid)
{
    if (q->cert_dropped.empty()) { q->tables_done = false; q->fk_configs_seen = 0; return CC_OK; }
    bool replan = false, any = false;
    for (size_t p = 0; p < q->cert_dropped.size(); ++p) {
        if (invalid && !(*invalid)[p]) continue;
        if (q->cert_dropped[p]) { q->cert_dropped[p] = 0; q->ndropped--; replan = true; any = true; }
        if (q->table_of_orig[p] >= 0) { q->table_of_orig[p] = -1; any = true; }
    }
    if (!any) return CC_OK;
    q->certs.erase(std::remove_if(q->certs.begin(), q->certs.end(), [&](const cc_certificate& c) { return !invalid || (*invalid)[c.pair]; }), q->certs.end());
    q->tables_done = false; q->fk_configs_seen = 0;
    CC_CUDA(cudaSetDevice(q->scene->device));
    CC_CUDA(cudaDeviceSynchronize());
    if (replan) {
        std::vector<int> keep;
        std::vector<uint8_t> main_dynamic(q->nq, 0);
        for (int p : q->p_main_all) {
            if (q->cert_dropped[p]) continue;
            keep.push_back(p);
            const int a = q->quads[4 * p];
            if (q->moves[a]) main_dynamic[a] = 1;
            if (q->quads[4 * p + 3] == 1) { const int b2 = q->quads[4 * p + 2]; if (q->moves[b2]) main_dynamic[b2] = 1; }
        }
        q->p_main = keep;
        int rc = build_plan(q, q->mainl, q->p_main, main_dynamic, true);
        if (rc != CC_OK) return rc;
    }
    return upload_pair_tables(q);
}

// Re-read what a bound query captured from the scene — the poses of every static slot (other bodies' links, query links no
// active DOF moves), the query body's base pose, the values of its inactive DOFs — without rebuilding the pair lists.
int cc_query_refresh(cc_query* q)
{
    if (!q) return fail(CC_ERR_INVALID, "query is NULL");
    cc_scene* S = q->scene;
    HBody* B = get_body(S, q->body);
    if (!B) return fail(CC_ERR_STATE, "the query's body was removed");
    if (!S->committed) return fail(CC_ERR_STATE, "scene changed since the query was bound");
    CC_CUDA(cudaSetDevice(S->device));
    CC_CUDA(cudaDeviceSynchronize());
    // what moved since the snapshot: the query body's base, and which other bodies (the static slots of the plan that holds
    // EVERY live pair — a body whose pairs are all certified no longer appears in the batched plan)
    const bool base_moved = memcmp(q->base, B->links[0].xf, 7 * sizeof(double)) != 0;
    std::vector<uint8_t> body_moved(S->bodies.size(), 0);
    for (int i = 0; i < q->full.nstatic && (size_t)i * 8 < q->full.st_pose.size(); ++i) {
        const std::pair<int, int>& sl = q->full.statics[i];
        if (memcmp(&q->full.st_pose[(size_t)i * 8], S->bodies[sl.first]->links[sl.second].xf, 7 * sizeof(double)) != 0) body_moved[sl.first] = 1;
    }
    memcpy(q->base, B->links[0].xf, sizeof(q->base));
    q->pose_epoch = S->pose_epoch;
    q->center[0] = q->base[4]; q->center[1] = q->base[5]; q->center[2] = q->base[6];
    for (int d = 0; d < B->ndof && d < (int)q->col_of_dof.size(); ++d) if (q->col_of_dof[d] >= 0) q->cur_active[q->col_of_dof[d]] = B->dofvals[d];
    // inactive DOF values are folded into the chain programs' constants: did any change?
    bool inactive_changed = false;
    for (ccdev::FkStep& st : q->steps) {
        if (st.src_col >= 0 || (st.flags & 1)) continue;
        for (const cc_joint& j : B->joints)
            if (j.parent_link == st.parent && j.child_link == st.child) {
                const int srcdof = j.mimic_dof >= 0 ? j.mimic_dof : j.dof_index;
                const double v = srcdof >= 0 ? B->dofvals[srcdof] : 0.0;
                if (v != st.src_const) { st.src_const = v; inactive_changed = true; }
                break;
            }
    }
    // free-space certificates were proven on the old poses: drop the ones a moved body touches — every environment pair when
    // the robot's base moved, everything when an inactive DOF changed (it may sit on the path between two links) — and keep
    // the rest; the dropped ones are re-examined after the next 16k configurations
    if (!q->cert_dropped.empty()) {
        std::vector<uint8_t> invalid(q->cert_dropped.size(), 0);
        for (size_t p = 0; p < invalid.size(); ++p) {
            const int ob = q->quads[4 * p + 1], cls = q->quads[4 * p + 3];
            invalid[p] = inactive_changed || (cls == 0 && (base_moved || body_moved[ob])) ? 1 : 0;
        }
        const int rc = invalidate_certificates(q, &invalid);
        if (rc != CC_OK) return rc;
    }
    Plan* plans[6] = {&q->full, &q->mainl, &q->hoisted, q->dist_ready ? &q->dist : nullptr, q->dist_main_ready ? &q->dist_main : nullptr,
                      q->dist_main_ready ? &q->dist_static : nullptr};
    for (auto& sm : q->dist_static_min) sm.valid = false;      // measured on the old poses
    for (Plan* pl : plans) {
        if (!pl || pl->nstatic == 0) continue;
        std::vector<float> stxf((size_t)pl->nstatic * 12, 0.f), staabb((size_t)pl->nstatic * 8, 0.f);
        for (int i = 0; i < pl->nstatic; ++i) {
            const HLink& L = S->bodies[pl->statics[i].first]->links[pl->statics[i].second];
            pose_to_rt32(L.xf, q->center, &stxf[(size_t)i * 12]);
            static_aabb(L, q->center, &staabb[(size_t)i * 8]);
            if ((size_t)i * 8 + 8 <= pl->st_pose.size()) memcpy(&pl->st_pose[(size_t)i * 8], L.xf, 8 * sizeof(double));
        }
        CC_CUDA(pl->d_st_xf.upload(stxf)); CC_CUDA(pl->d_st_aabb.upload(staabb));
    }
    if (inactive_changed)
        for (Plan* pl : plans) {
            if (!pl || pl->steps.empty()) continue;
            compile_fk_program(q, *pl);
            CC_CUDA(pl->d_steps.upload(pl->steps));
        }
    CC_CUDA(cudaDeviceSynchronize());
    {
        const int rc = evaluate_hoisted(q);
        if (rc != CC_OK) return rc;
    }
    for (cc_query* pq : q->peers) { const int rc = cc_query_refresh(pq); if (rc != CC_OK) return rc; }
    if (!q->peers.empty()) CC_CUDA(cudaSetDevice(S->device));
    return CC_OK;
}

int cc_query_destroy(cc_query* q)
{
    if (!q) return CC_OK;
    for (cc_query* pq : q->peers) cc_query_destroy(pq);
    cudaSetDevice(q->scene->device);
    cudaDeviceSynchronize();
    delete q;
    return CC_OK;
}

int cc_query_get_info(const cc_query* q, cc_query_info* info)
{
    if (!q || !info) return fail(CC_ERR_INVALID, "bad arguments");
    info->ndof_active = q->ncols;
    info->nexternal = q->nexternal;
    info->reserved = 0;
    info->nlinks = q->nq;
    info->npairs_env = q->nenv_total;
    info->npairs_self = q->nself_total;
    info->npairs_hoisted = (int)q->hoisted.pairs.size();
    info->mask_words = q->mask_words;
    info->total_nodes = q->scene->total_nodes;
    info->total_tris = q->scene->total_tris;
    return CC_OK;
}

int cc_query_get_pairs(const cc_query* q, int32_t* out)
{
    if (!q || !out) return fail(CC_ERR_INVALID, "bad arguments");
    if (!q->quads.empty()) memcpy(out, q->quads.data(), q->quads.size() * sizeof(int32_t));
    return CC_OK;
}

int cc_query_get_links(const cc_query* q, int32_t* out)
{
    if (!q || !out) return fail(CC_ERR_INVALID, "bad arguments");
    for (int l = 0; l < q->nq; ++l) { out[2 * l] = q->ql_body[l]; out[2 * l + 1] = q->ql_link[l]; }
    return CC_OK;
}

int cc_query_trace_work(cc_query* q, int enabled)
{
    if (!q) return fail(CC_ERR_INVALID, "query is NULL");
    q->trace_work = enabled != 0;
    return CC_OK;
}

int cc_query_get_work_trace(cc_query* q, uint32_t* out, int n)
{
    if (!q || !out || n < 0 || n > q->work_n) return fail(CC_ERR_INVALID, "bad arguments");
    CC_CUDA(cudaSetDevice(q->scene->device));
    CC_CUDA(cudaDeviceSynchronize());
    if (n) CC_CUDA(cudaMemcpy(out, q->d_work.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return CC_OK;
}

int cc_check_batch_device(cc_query* q, const double* d_q, int n, int flags, uint32_t* d_verdict, int32_t* d_fe,
                          int32_t* d_fs, uint32_t* d_mask, void* stream)
{
    if (!q || n < 0 || (n > 0 && !d_q && q->ncols > 0)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!q->has_fk) return fail(CC_ERR_STATE, "body has no usable kinematics: use cc_check_poses");
    if (!q->scene->committed) return fail(CC_ERR_STATE, "scene changed since the query was bound");
    CC_CUDA(cudaSetDevice(q->scene->device));
    const int f = (flags & (CC_BATCH_ENV | CC_BATCH_SELF | CC_BATCH_ALL_PAIRS)) | ((flags & CC_BATCH_FIRST_HIT) ? ccdev::kFlagFirstHit : 0);
    { int rc = maybe_build_tables(q, n); if (rc != CC_OK) return rc; }
    return launch(q, q->mainl, true, d_q, n, f, d_verdict, d_fe, d_fs, d_mask, nullptr, (cudaStream_t)stream);
}

static int run_host(cc_query* q, const double* in, size_t in_doubles, int n, int kflags, bool pose_mode,
                    uint32_t* verdict, int32_t* fe, int32_t* fs, uint32_t* mask, double* fk)
{
    cc_scene* S = q->scene;
    if (n == 0) return CC_OK;
    CC_CUDA(cudaSetDevice(S->device));
    CC_CUDA(q->d_q.reserve(std::max<size_t>(1, in_doubles)));
    if (verdict) CC_CUDA(q->d_verdict.reserve(n));
    if (fe) CC_CUDA(q->d_fe.reserve(n));
    if (fs) CC_CUDA(q->d_fs.reserve(n));
    if (mask) CC_CUDA(q->d_mask.reserve((size_t)n * q->mask_words));
    if (fk) CC_CUDA(q->d_fk.reserve((size_t)n * q->nq * 8));
    const Plan& pl = pose_mode ? q->full : (fk ? q->hoisted : q->mainl);   // FK read-back needs every joint: the unpruned program
    HostIO hio;
    hio.in = in_doubles ? in : nullptr; hio.verdict = verdict; hio.fe = fe; hio.fs = fs; hio.mask = mask; hio.fk = fk;
    int rc = launch(q, pl, !pose_mode, q->d_q.p, n, kflags, verdict ? q->d_verdict.p : nullptr, fe ? q->d_fe.p : nullptr,
                    fs ? q->d_fs.p : nullptr, mask ? q->d_mask.p : nullptr, fk ? q->d_fk.p : nullptr, 0, &hio);
    if (rc != CC_OK) return rc;
    CC_CUDA(cudaStreamSynchronize(0));
    return check_err_flag(S);
}

// ---- multi-device scenes: a host-buffer batch is range-sharded over the replicas, one host thread per extra device, every
// replica copying its slice in and its results out of the caller's buffers itself (host consumer, no collective)
static int shard_count(const cc_query* q, int n)
{
    const int G = 1 + (int)q->peers.size();
    if (G == 1) return 1;
    static const int min_per = getenv("CC_TUNE_SHARD_MIN") ? std::max(1, atoi(getenv("CC_TUNE_SHARD_MIN"))) : 16384;
    return std::max(1, std::min(G, n / min_per));        // small batches stay on the first device: a launch costs more than the work
}

}  // extern "C"
// f(replica query, first row, rows) -> cc_status; shard 0 runs on the calling thread
template <typename F>
static int run_sharded(cc_query* q, int n, F f)
{
    const int G = shard_count(q, n);
    if (G <= 1) return f(q, 0, n);
    std::vector<int> rcs(G, CC_OK);
    std::vector<std::string> errs(G);
    auto begin = [&](int g) { return (int)((long long)n * g / G); };
    std::vector<std::thread> workers;
    for (int g = 1; g < G; ++g)
        workers.emplace_back([&, g]() {
            rcs[g] = f(q->peers[g - 1], begin(g), begin(g + 1) - begin(g));
            if (rcs[g] != CC_OK) errs[g] = g_err;
        });
    rcs[0] = f(q, 0, begin(1));
    if (rcs[0] != CC_OK) errs[0] = g_err;
    for (std::thread& t : workers) t.join();
    for (int g = 0; g < G; ++g)
        if (rcs[g] != CC_OK) { g_err = "device " + std::to_string(g == 0 ? q->scene->device : q->peers[g - 1]->scene->device) + ": " + errs[g]; return rcs[g]; }
    return CC_OK;
}
extern "C" {

int cc_check_batch(cc_query* q, const double* dof_values, int n, int flags, uint32_t* verdict, int32_t* first_env,
                   int32_t* first_self, uint32_t* pair_mask)
{
    if (!q || n < 0 || (n > 0 && q->ncols > 0 && !dof_values)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!q->has_fk) return fail(CC_ERR_STATE, "body has no usable kinematics: use cc_check_poses");
    if (!q->scene->committed) return fail(CC_ERR_STATE, "scene changed since the query was bound");
    const int f = (flags & (CC_BATCH_ENV | CC_BATCH_SELF | CC_BATCH_ALL_PAIRS)) | ((flags & CC_BATCH_FIRST_HIT) ? ccdev::kFlagFirstHit : 0);
    return run_sharded(q, n, [&](cc_query* r, int off, int m) -> int {
        { int rc = maybe_build_tables(r, m); if (rc != CC_OK) return rc; }
        return run_host(r, dof_values ? dof_values + (size_t)off * r->ncols : nullptr, (size_t)m * r->ncols, m, f, false,
                        verdict ? verdict + off : nullptr, first_env ? first_env + off : nullptr, first_self ? first_self + off : nullptr,
                        pair_mask ? pair_mask + (size_t)off * r->mask_words : nullptr, nullptr);
    });
}

int cc_check_poses(cc_query* q, const double* T8, int n, int flags, uint32_t* verdict, int32_t* first_env,
                   int32_t* first_self, uint32_t* pair_mask)
{
    if (!q || n < 0 || (n > 0 && !T8)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!q->scene->committed) return fail(CC_ERR_STATE, "scene changed since the query was bound");
    const int f = (flags & (CC_BATCH_ENV | CC_BATCH_SELF | CC_BATCH_ALL_PAIRS)) | ccdev::kFlagPoseMode | ((flags & CC_BATCH_FIRST_HIT) ? ccdev::kFlagFirstHit : 0);
    return run_host(q, T8, (size_t)n * q->nq * 8, n, f, true, verdict, first_env, first_self, pair_mask, nullptr);
}

int cc_fk_batch(cc_query* q, const double* dof_values, int n, double* out_T8)
{
    if (!q || n < 0 || !out_T8 || (n > 0 && q->ncols > 0 && !dof_values)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!q->has_fk) return fail(CC_ERR_STATE, "body has no usable kinematics");
    return run_sharded(q, n, [&](cc_query* r, int off, int m) -> int {
        return run_host(r, dof_values ? dof_values + (size_t)off * r->ncols : nullptr, (size_t)m * r->ncols, m, ccdev::kFlagFkOnly, false,
                        nullptr, nullptr, nullptr, nullptr, out_T8 + (size_t)off * r->nq * 8);
    });
}

int cc_distance_batch(cc_query* q, const double* dof_values, int n, int flags, double max_distance, double rel_err,
                      double abs_err, double* min_distance, int32_t* min_pair, double* closest_points)
{
    if (!q || n < 0 || !min_distance || (n > 0 && q->ncols > 0 && !dof_values)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!q->has_fk) return fail(CC_ERR_STATE, "body has no usable kinematics: use cc_distance_poses");
    if (!q->scene->committed) return fail(CC_ERR_STATE, "scene changed since the query was bound");
    return run_sharded(q, n, [&](cc_query* r, int off, int m) -> int {
        if (!r->dist_main_ready) {
            CC_CUDA(cudaSetDevice(r->scene->device));
            std::vector<uint8_t> all_dynamic(r->nq, 1);
            int rc = build_plan(r, r->dist_main, r->p_dist_main, r->dist_dynamic, true);
            if (rc == CC_OK) rc = build_plan(r, r->dist_static, r->p_dist_static, all_dynamic, true);
            if (rc != CC_OK) return rc;
            CC_CUDA(cudaDeviceSynchronize());
            r->dist_main_ready = true;
        }
        // the pairs no active DOF can change: measured once per pose snapshot and class selection (exact, unbounded)
        cc_query::StaticMin& sm = r->dist_static_min[flags & 3];
        static const bool tune_no_static = getenv("CC_TUNE_DIST_STATIC") && atoi(getenv("CC_TUNE_DIST_STATIC")) == 0;      // A/B knob
        if (!sm.valid) {
            sm.d = 1e30; sm.pair = -1;
            if (!r->dist_static.pairs.empty()) {
                const int rc = launch_distance(r, r->dist_static, false, r->cur_active.data(), 1, flags, 0.0, 0.0, 0.0, &sm.d, &sm.pair, sm.pts);
                if (rc != CC_OK) return rc;
            }
            sm.valid = true;
        }
        const bool have = sm.pair >= 0 && (max_distance <= 0 || sm.d < max_distance);
        double* dist = min_distance + off;
        int32_t* pair = min_pair ? min_pair + off : nullptr;
        double* pts = closest_points ? closest_points + (size_t)off * 6 : nullptr;
        std::vector<int32_t> pair_tmp;
        if (have && !pair) { pair_tmp.resize(m); pair = pair_tmp.data(); }       // needed to tell "nothing closer" apart
        if (r->dist_main.pairs.empty() || (have && sm.d <= 0.0)) {
            for (int i = 0; i < m; ++i) { dist[i] = max_distance > 0 ? max_distance : 1e30; if (pair) pair[i] = -1; if (pts) for (int k = 0; k < 6; ++k) pts[6 * (size_t)i + k] = 0.0; }
        } else {
            const int rc = launch_distance(r, r->dist_main, false, dof_values ? dof_values + (size_t)off * r->ncols : nullptr, m, flags, max_distance,
                                           rel_err, abs_err, dist, pair, pts, false, (have && !tune_no_static) ? sm.d : -1.0);
            if (rc != CC_OK) return rc;
        }
        if (have)
            for (int i = 0; i < m; ++i)
                if (pair[i] < 0 || sm.d < dist[i]) {
                    dist[i] = sm.d; pair[i] = sm.pair;
                    if (pts) memcpy(pts + 6 * (size_t)i, sm.pts, sizeof(sm.pts));
                }
        return CC_OK;
    });
}

int cc_distance_poses(cc_query* q, const double* T8, int n, int flags, double max_distance, double rel_err, double abs_err,
                      double* min_distance, int32_t* min_pair, double* closest_points)
{
    if (!q || n < 0 || !min_distance || (n > 0 && !T8)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!q->scene->committed) return fail(CC_ERR_STATE, "scene changed since the query was bound");
    return launch_distance(q, q->full, true, T8, n, flags, max_distance, rel_err, abs_err, min_distance, min_pair, closest_points);
}

}  // extern "C" (reopened below)
#include "cc_contacts.inl"
extern "C" {

// CO_UseTolerance: which pairs are within `tolerance` of each other (PQP_Tolerance: distance <= tolerance, pqp/PQP.cpp:1056).
// Per-pair mode of the distance kernel with the search radius just above the tolerance; one bit per pair of the query.
static int tolerance_common(cc_query* q, const Plan& pl, bool pose_mode, const double* in, int n, int flags, double tolerance,
                            uint32_t* within_mask)
{
    if (n == 0) return CC_OK;
    const int npd = (int)pl.pairs.size();
    memset(within_mask, 0, (size_t)n * q->mask_words * sizeof(uint32_t));
    if (npd == 0) return CC_OK;
    std::vector<double> dist((size_t)n * npd);
    int rc = launch_distance(q, pl, pose_mode, in, n, flags, tolerance * 1.001 + 1e-6, 0.0, 0.0, dist.data(), nullptr, nullptr, true);
    if (rc != CC_OK) return rc;
    const bool env = (flags & CC_BATCH_ENV) != 0, self = (flags & CC_BATCH_SELF) != 0;
    for (int c = 0; c < n; ++c)
        for (int k = 0; k < npd; ++k) {
            const bool is_env = k < pl.nenv;
            if ((is_env && !env) || (!is_env && !self)) continue;
            if (dist[(size_t)c * npd + k] <= tolerance) {
                const int p = pl.orig[k];
                within_mask[(size_t)c * q->mask_words + (p >> 5)] |= 1u << (p & 31);
            }
        }
    return CC_OK;
}

int cc_tolerance_batch(cc_query* q, const double* dof_values, int n, int flags, double tolerance, uint32_t* within_mask)
{
    if (!q || n < 0 || !within_mask || !(tolerance >= 0) || (n > 0 && q->ncols > 0 && !dof_values)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!q->has_fk) return fail(CC_ERR_STATE, "body has no usable kinematics: use cc_tolerance_poses");
    if (!q->scene->committed) return fail(CC_ERR_STATE, "scene changed since the query was bound");
    if (!q->dist_ready) {
        CC_CUDA(cudaSetDevice(q->scene->device));
        std::vector<uint8_t> all_dynamic(q->nq, 1);
        int rc = build_plan(q, q->dist, q->p_full, all_dynamic, true);
        if (rc != CC_OK) return rc;
        CC_CUDA(cudaDeviceSynchronize());
        q->dist_ready = true;
    }
    return tolerance_common(q, q->dist, false, dof_values, n, flags, tolerance, within_mask);
}

int cc_tolerance_poses(cc_query* q, const double* T8, int n, int flags, double tolerance, uint32_t* within_mask)
{
    if (!q || n < 0 || !within_mask || !(tolerance >= 0) || (n > 0 && !T8)) return fail(CC_ERR_INVALID, "bad arguments");
    if (!q->scene->committed) return fail(CC_ERR_STATE, "scene changed since the query was bound");
    return tolerance_common(q, q->full, true, T8, n, flags, tolerance, within_mask);
}

int cc_host_alloc(size_t bytes, void** out)
{
    if (!out) return fail(CC_ERR_INVALID, "out is NULL");
    *out = nullptr;
    CC_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return CC_OK;
}

int cc_host_free(void* p)
{
    if (p) CC_CUDA(cudaFreeHost(p));
    return CC_OK;
}

int cc_scene_synchronize(cc_scene* s)
{
    if (!s) return fail(CC_ERR_INVALID, "scene is NULL");
    for (cc_scene* p : s->peers) { const int rc = cc_scene_synchronize(p); if (rc != CC_OK) return rc; }
    CC_CUDA(cudaSetDevice(s->device));
    CC_CUDA(cudaDeviceSynchronize());
    return check_err_flag(s);
}

int cc_scene_trace_timeline(cc_scene* s, int enabled)
{
    if (!s) return fail(CC_ERR_INVALID, "scene is NULL");
    s->trace_timeline = enabled != 0;
    if (!enabled) s->tl_lane.clear();
    return CC_OK;
}

int cc_scene_get_timeline(cc_scene* s, double* out_ms, int32_t* out_lane, int max_chunks, int* out_chunks)
{
    if (!s || !out_chunks || max_chunks < 0 || (max_chunks > 0 && !out_ms)) return fail(CC_ERR_INVALID, "bad arguments");
    CC_CUDA(cudaSetDevice(s->device));
    CC_CUDA(cudaDeviceSynchronize());
    const int nc = (int)s->tl_lane.size();
    *out_chunks = nc;
    for (int c = 0; c < nc && c < max_chunks; ++c) {
        for (int e = 0; e < 3; ++e) {
            float ms = 0.f;
            CC_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->tl_events[3 * c + e]));
            out_ms[3 * c + e] = (double)ms;
        }
        if (out_lane) out_lane[c] = s->tl_lane[c];
    }
    return CC_OK;
}

int cc_scene_set_counters(cc_scene* s, int enabled)
{
    if (!s) return fail(CC_ERR_INVALID, "scene is NULL");
    s->counters = enabled != 0;
    return forward_peers(s, [&](cc_scene* p) { return cc_scene_set_counters(p, enabled); });
}

int cc_device_fp32_peak(int cuda_device, double* out_tflops)
{
    if (!out_tflops) return fail(CC_ERR_INVALID, "out is NULL");
    CC_CUDA(cudaSetDevice(cuda_device));
    cudaDeviceProp prop;
    CC_CUDA(cudaGetDeviceProperties(&prop, cuda_device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    float* d = nullptr;
    CC_CUDA(cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    CC_CUDA(cudaEventCreate(&e0)); CC_CUDA(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CC_CUDA(cudaEventRecord(e0, 0));
        ccdev::cc_fma_peak_kernel<<<blocks, threads>>>(d, 0.999f, 1e-3f);
        CC_CUDA(cudaEventRecord(e1, 0));
        CC_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8 * 4096 * (double)blocks * threads;
        if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    *out_tflops = best;
    return CC_OK;
}

int cc_scene_set_pair_tables(cc_scene* s, int enabled)
{
    if (!s) return fail(CC_ERR_INVALID, "scene is NULL");
    s->pair_tables = enabled != 0;
    return forward_peers(s, [&](cc_scene* p) { return cc_scene_set_pair_tables(p, enabled); });
}

int cc_query_get_table_count(const cc_query* q, int* out_tables, int* out_dropped, double* out_build_ms)
{
    if (!q) return fail(CC_ERR_INVALID, "bad arguments");
    if (out_tables) *out_tables = q->ntables > 0 ? q->ntables : 0;
    if (out_dropped) *out_dropped = q->ndropped;
    if (out_build_ms) *out_build_ms = q->cert_build_ms;
    return CC_OK;
}

int cc_query_get_certificates(const cc_query* q, cc_certificate* out, int max_records, int* out_count)
{
    if (!q || !out_count || max_records < 0 || (max_records > 0 && !out)) return fail(CC_ERR_INVALID, "bad arguments");
    *out_count = (int)q->certs.size();
    for (int i = 0; i < max_records && i < (int)q->certs.size(); ++i) out[i] = q->certs[i];
    return CC_OK;
}

int cc_scene_set_tri_epsilon(cc_scene* s, double eps)
{
    if (!s || eps < 0) return fail(CC_ERR_INVALID, "bad arguments");
    s->tri_eps = (float)eps;
    return forward_peers(s, [&](cc_scene* p) { return cc_scene_set_tri_epsilon(p, eps); });
}

int cc_scene_get_stats(const cc_scene* sc, cc_stats* out)
{
    cc_scene* s = const_cast<cc_scene*>(sc);
    if (!s || !out) return fail(CC_ERR_INVALID, "bad arguments");
    CC_CUDA(cudaSetDevice(s->device));
    if (s->ev_pending) {
        CC_CUDA(cudaEventSynchronize(s->ev1));
        float ms = 0;
        CC_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->stats.last_kernel_ms = ms;
        s->ev_pending = false;
    }
    unsigned long long h[2] = {0, 0};
    CC_CUDA(cudaMemcpy(h, s->d_stats.p, sizeof(h), cudaMemcpyDeviceToHost));
    s->stats.bv_tests = h[0]; s->stats.tri_tests = h[1];
    *out = s->stats;
    for (cc_scene* p : s->peers) {       // a multi-device scene reports the sum over its replicas
        cc_stats ps;
        const int rc = cc_scene_get_stats(p, &ps);
        if (rc != CC_OK) return rc;
        out->configs += ps.configs; out->bv_tests += ps.bv_tests; out->tri_tests += ps.tri_tests; out->launches += ps.launches;
        out->last_kernel_ms = std::max(out->last_kernel_ms, ps.last_kernel_ms);
    }
    if (!s->peers.empty()) CC_CUDA(cudaSetDevice(s->device));
    return CC_OK;
}

}  // extern "C"

// Diagnostic, no device needed: the broadphase kernel's box test on the host — the very functions cc_broad_kernel compiles
// (quantise_box, qbox_overlap), so that a CPU test can check that an overlap of the FP32 boxes is never lost.
extern "C" int cc_debug_motion_classes(const cc_joint* joints, int njoints, int nlinks, int ndof, const uint8_t* affects, const int32_t* active,
                                       int nactive, int base_moves, const int32_t* pairs, int npairs, uint8_t* out_moves, uint8_t* out_invariant)
{
    if (!joints || njoints < 0 || nlinks <= 0 || ndof < 0 || (ndof && !affects) || nactive < 0 || (nactive && !active) || npairs < 0 ||
        (npairs && (!pairs || !out_invariant)) || !out_moves)
        return fail(CC_ERR_INVALID, "cc_debug_motion_classes: bad arguments");
    std::vector<cc_joint> J(joints, joints + njoints);
    std::vector<int> act(active, active + nactive), col_of_dof((size_t)ndof, -1);
    for (int c = 0; c < nactive; ++c) {
        if (act[c] < 0 || act[c] >= ndof) return fail(CC_ERR_INVALID, "cc_debug_motion_classes: active DOF index out of range");
        col_of_dof[act[c]] = c;
    }
    MotionClasses m;
    m.build(J, nlinks, act, col_of_dof, base_moves != 0, [&](int d, int l) { return affects[(size_t)d * nlinks + l] != 0; });
    memcpy(out_moves, m.moves.data(), (size_t)nlinks);
    for (int p = 0; p < npairs; ++p) {
        const int a = pairs[2 * p], b = pairs[2 * p + 1];
        if (a < 0 || a >= nlinks || b < 0 || b >= nlinks) return fail(CC_ERR_INVALID, "cc_debug_motion_classes: link out of range");
        out_invariant[p] = m.pose_invariant(a, b) ? 1 : 0;
    }
    return CC_OK;
}

extern "C" int cc_debug_broad_overlap(const float* boxes_a, const float* boxes_b, int n, uint8_t* out)
{
    if (!boxes_a || !boxes_b || !out || n < 0) return fail(CC_ERR_INVALID, "cc_debug_broad_overlap: bad arguments");
    for (int i = 0; i < n; ++i) {
        const float* a = boxes_a + 8 * (size_t)i;
        const float* b = boxes_b + 8 * (size_t)i;
        ccdev::Box8 A, B;
        A.c = make_float4(a[0], a[1], a[2], a[3]); A.e = make_float4(a[4], a[5], a[6], a[7]);
        B.c = make_float4(b[0], b[1], b[2], b[3]); B.e = make_float4(b[4], b[5], b[6], b[7]);
        bool exact = true;
        for (int k = 0; k < 3; ++k) exact = exact && fabs((double)a[k] - (double)b[k]) <= (double)a[4 + k] + (double)b[4 + k];
        const bool quantised = ccdev::qbox_overlap(ccdev::quantise_box(A), ccdev::quantise_box(B));
        out[i] = (uint8_t)((exact ? 1 : 0) | (quantised ? 2 : 0));
    }
    return CC_OK;
}

// Diagnostic, no device needed: the BVH of one triangle mesh exactly as cc_scene_commit builds it (binned surface-area
// splits, one triangle per leaf, children adjacent).  The CPU tests check its invariants; the GPU never sees this call.
extern "C" int cc_debug_build_bvh(const double* vertices, int nvertices, const int32_t* triangles, int ntriangles, float* out_nodes,
                                  int max_nodes, int* out_count)
{
    if (!vertices || !triangles || nvertices <= 0 || ntriangles <= 0 || !out_count) return fail(CC_ERR_INVALID, "cc_debug_build_bvh: bad arguments");
    for (int i = 0; i < 3 * ntriangles; ++i)
        if (triangles[i] < 0 || triangles[i] >= nvertices) return fail(CC_ERR_INVALID, "cc_debug_build_bvh: vertex index out of range");
    HLink L;
    L.verts.assign(vertices, vertices + 3 * (size_t)nvertices);
    L.tris.assign(triangles, triangles + 3 * (size_t)ntriangles);
    build_link_bvh(L);
    *out_count = (int)L.nodes.size();
    if (out_nodes && max_nodes > 0)
        memcpy(out_nodes, L.nodes.data(), sizeof(HNode) * (size_t)std::min<int>(max_nodes, (int)L.nodes.size()));
    return CC_OK;
}

