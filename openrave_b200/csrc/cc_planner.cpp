// cc_planner.cpp — BiRRT + linear shortcutting on the batched collision entry (host side, C++20).
//
// What the reference does (paths relative to the OpenRAVE tree) and what this file restates:
//   plugins/rplanners/rrt.h:279-368   BirrtPlanner::InitPlan        start/goal validation, tree roots
//   plugins/rplanners/rrt.h:370-540   BirrtPlanner::PlanPath        sample / extend A / extend B toward A / swap
//   plugins/rplanners/rplanners.h:308-460  SpatialTree::Extend      step-length ray toward the target, <= 100 steps
//   plugins/rplanners/rrt.h:151-240   RrtPlanner::_SimpleOptimizePath   10 random shortcuts on the raw path
//   plugins/rplanners/rrt.h:617-659   BirrtPlanner::_ExtractPath
//   plugins/rplanners/linearsmoother.cpp:279-383, 385-559, 1235-1263   LinearSmoother (default post-processor)
//   src/libopenrave/planningutils.cpp:2586-3629  DynamicsCollisionConstraint::Check, linear branch (the sample sequence)
//   src/libopenrave/planner.cpp:88-112, 650-697  AddStatesWithLimitCheck, SetRobotActiveJoints
//   plugins/basesamplers/mt19937ar.h:101-162, robotconfiguration.cpp:81-112   the samplers
//
// What is different by design: the reference evaluates one configuration at a time and stops at the first failure.  Here
// every planning problem is a coroutine that suspends whenever it needs a verdict; the driver collects the pending request
// of every live problem (a "group" of ordered segments), sends all of them through ONE cc_check_batch call, and resumes
// the problems with the index of their first failing segment.  The configurations of an Extend ray, a shortcut or a
// smoother candidate never depend on a collision result inside the same request, so the outcome equals the serial loop's.
// The nearest-neighbour structure is an exact linear scan (the reference's cover tree, rplanners.h:170-860, returns the
// same node up to ties); insertion refuses a node closer than 0.001 * step to its nearest neighbour like _InsertRecursive.
#include "cc_epsilons.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <coroutine>
#include <cstdint>
#include <cstring>
#include <list>
#include <memory>
#include <random>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cudacollision.h"

extern "C" void cc_internal_set_error(const char* msg);   // cudacollision.cu (thread-local cc_last_error text)

namespace {

int fail(int code, const char* msg) { cc_internal_set_error(msg); return code; }

constexpr double kPi = 3.14159265358979323846;
constexpr double kEpsilon = CC_EPSILON;                          // g_fEpsilon for dReal = double (openrave.h:130)
constexpr double kEpsilonLinear = CC_EPSILON_LINEAR;             // libopenrave.h:140
constexpr double kEpsilonJointLimit = CC_EPSILON_JOINT_LIMIT;    // libopenrave.h:141
constexpr double kEpsilonDotProduct = CC_EPSILON_JOINT_LIMIT;    // RavePow(g_fEpsilon, 0.8), rrt.h:22
enum Interval { IT_Open = 0, IT_OpenStart = 1, IT_OpenEnd = 2, IT_Closed = 3 };

// ------------------------------------------------------------------------------------------------ configuration space
struct CSpace {
    int n = 0;
    std::vector<double> lo, hi, res, w2;
    std::vector<uint8_t> circ;
    bool any_circular = false;

    // KinBody::SubtractDOFValues (kinbody.cpp:1357-1394): a <- a - b, circular DOFs wrapped into [-pi, pi)
    void diff(double* a, const double* b) const
    {
        for (int i = 0; i < n; ++i) {
            double d = a[i] - b[i];
            if (any_circular && circ[i]) {
                if (d < -kPi) d += 2 * kPi * std::ceil((-kPi - d) / (2 * kPi));
                else if (d >= kPi) d -= 2 * kPi * std::floor((d + kPi) / (2 * kPi));
            }
            a[i] = d;
        }
    }
    // SimpleDistanceMetric::Eval (planningutils.cpp:4714-4723)
    double dist(const double* c0, const double* c1) const
    {
        double s = 0;
        for (int i = 0; i < n; ++i) {
            double d = c0[i] - c1[i];
            if (any_circular && circ[i]) {
                if (d < -kPi) d += 2 * kPi * std::ceil((-kPi - d) / (2 * kPi));
                else if (d >= kPi) d -= 2 * kPi * std::floor((d + kPi) / (2 * kPi));
            }
            s += w2[i] * d * d;
        }
        return std::sqrt(s);
    }
    // AddStatesWithLimitCheck (planner.cpp:88-112): returns true when the clamp moved q by more than the joint epsilon
    bool neigh(double* q, const double* delta) const
    {
        bool deviated = false;
        for (int i = 0; i < n; ++i) {
            q[i] += delta[i];
            if (q[i] > hi[i]) { if (q[i] > hi[i] + kEpsilonJointLimit) deviated = true; q[i] = hi[i]; }
            else if (q[i] < lo[i]) { if (q[i] < lo[i] - kEpsilonJointLimit) deviated = true; q[i] = lo[i]; }
        }
        return deviated;
    }
};

// mt19937 sampler (plugins/basesamplers/mt19937ar.h): same generator as std::mt19937 seeded with init_genrand(seed); the
// float constants 1.0f/4294967295.0f and 1.0f/4294967296.0f both round to 2^-32.
struct Sampler {
    std::mt19937 gen;
    explicit Sampler(uint32_t seed = 0) : gen(seed) {}
    uint32_t u32() { return (uint32_t)gen(); }
    double real(Interval it = IT_Closed)
    {
        const double k = 1.0 / 4294967296.0;
        const double r = (double)u32();
        if (it == IT_Open) return (r + 0.5) * k;
        if (it == IT_OpenStart) return (r + 1.0) * k;
        return r * k;
    }
};

// ------------------------------------------------------------------------------------------------ edge sample sequence
// DynamicsCollisionConstraint::Check, positions-only linear branch.  Order: q1 first when the interval is closed at the
// end (:2692-2708), then q0 when closed at the start (:2809-2826), then steps start..numSteps-1 produced by repeated
// _neighstatefn from the previous sample (:3405-3454), then the residual sample when the walk did not land on q1 (:3524).
// Returns the number of appended configurations or -1 when a step leaves the joint limits.
int edge_samples(const CSpace& cs, const double* q0, const double* q1, int interval, std::vector<double>& out,
                 std::vector<double>* times = nullptr)
{
    const int n = cs.n;
    int start = (interval == IT_Open || interval == IT_OpenStart) ? 1 : 0;
    const bool check_end = (interval == IT_OpenStart || interval == IT_Closed);
    int count = 0;
    auto push = [&](const double* q, double t) { out.insert(out.end(), q, q + n); if (times) times->push_back(t); ++count; };
    if (check_end) push(q1, 1.0);
    static thread_local std::vector<double> dQ, cur, d;     // scratch (this runs a hundred times per Extend ray)
    dQ.assign(q1, q1 + n);
    cs.diff(dQ.data(), q0);
    int nsteps = 0, total = 0;
    for (int i = 0; i < n; ++i) {
        int s = cs.res[i] != 0 ? (int)(std::fabs(dQ[i]) / cs.res[i] + 0.99) : (int)(std::fabs(dQ[i]) * 100);
        total += s;
        nsteps = std::max(nsteps, s);
    }
    if (total == 0 && start > 0) return count;
    if (start == 0) { push(q0, 0.0); start = 1; }
    if (nsteps == 0) return count;
    const double fisteps = 1.0 / nsteps;
    for (int i = 0; i < n; ++i) dQ[i] *= fisteps;
    cur.assign(q0, q0 + n);
    d.resize(n);
    if (cs.neigh(cur.data(), dQ.data())) return -1;
    for (int f = start; f < nsteps; ++f) {
        push(cur.data(), f * fisteps);
        double scale = 1.0;
        bool big = false;
        for (int i = 0; i < n; ++i) {
            d[i] = q0[i] + (f + 1) * dQ[i] - cur[i];
            if (std::fabs(d[i]) > cs.res[i]) { scale = big ? std::min(scale, cs.res[i] / std::fabs(d[i])) : cs.res[i] / std::fabs(d[i]); big = true; }
        }
        if (big) for (int i = 0; i < n; ++i) d[i] *= scale;
        if (cs.neigh(cur.data(), d.data())) return -1;
    }
    double e = 0;
    for (int i = 0; i < n; ++i) e += (cur[i] - q1[i]) * (cur[i] - q1[i]);
    if (std::sqrt(e) > 1e-7) push(cur.data(), 1.0);
    return count;
}

// mathextra::solvequad (include/openrave/mathextra.h:214-256): roots of a x^2 + b x + c, the small-root forms kept
int solvequad(double a, double b, double c, double& r1, double& r2)
{
    if (a == 0) {
        if (b == 0) return 0;
        r1 = r2 = -c / b;
        return 1;
    }
    if (c == 0) { r1 = 0; r2 = -b / a; return 2; }
    double d = b * b - 4.0 * c * a;
    if (d < 0) {
        if (d < -1e-16) return 0;
        d = 0;
    }
    d = std::sqrt(d);
    if (std::fabs(-b - d) < std::fabs(a)) r1 = 0.5 * (-b + d) / a; else r1 = 2.0 * c / (-b - d);
    if (std::fabs(-b + d) < std::fabs(a)) r2 = 0.5 * (-b - d) / a; else r2 = 2.0 * c / (-b + d);
    return 2;
}

// DynamicsCollisionConstraint::Check, quadratic branch (planningutils.cpp:2726-2766, 2844-3319): the edge is the ramp
// q(t) = q0 + dq0 t + a t^2 / 2 with a = (dq1 - dq0) / timeelapsed.  The DOF that needs the most steps (counting both sides of
// its turning point) is advanced by equal increments of its own travelled distance; the time of every increment comes from
// the ramp's quadratic, and a step in which another DOF would move farther than 1.01 x its resolution is cut short in time.
// Same order as the linear branch: q1 first when the interval is closed at the end, q0 when closed at the start, then the
// walk.  `filled` holds the cc_ramp_sample_kind of every configuration: only kind 1 is reported by CFO_FillCheckedConfiguration
// (linear sub-steps after a step that still exceeds 1.01 x a resolution are checked but not reported, :3163-3196, 3207-3258).  Returns the number of appended configurations or -1 = CFO_StateSettingError.
int edge_samples_quadratic(const CSpace& cs, const double* q0, const double* q1, const double* dq0, const double* dq1, double T,
                           int interval, std::vector<double>& out, std::vector<double>* times, std::vector<uint8_t>* filled)
{
    const int n = cs.n;
    int start = (interval == IT_Open || interval == IT_OpenStart) ? 1 : 0;
    const bool check_end = (interval == IT_OpenStart || interval == IT_Closed);
    int count = 0;
    auto push = [&](const double* q, double t, int f) {
        out.insert(out.end(), q, q + n);
        if (times) times->push_back(t);
        if (filled) filled->push_back((uint8_t)f);
        ++count;
    };
    if (check_end) push(q1, T, 1);
    std::vector<double> dQ(q1, q1 + n), acc(n), cur(q0, q0 + n), vel(dq0, dq0 + n), prev(n), prevvel(n);
    cs.diff(dQ.data(), q0);
    const double iT = 1.0 / T;
    for (int i = 0; i < n; ++i) acc[i] = (dq1[i] - dq0[i]) * iT;
    int nsteps = 0, largest = -1, total = 0;
    for (int i = 0; i < n; ++i) {
        double travel = std::fabs(dQ[i]);
        if (std::fabs(acc[i]) > kEpsilonLinear) {
            const double tin = -dq0[i] / acc[i];
            if (tin >= 0 && tin < T) {                       // the DOF turns around inside the ramp: count both legs
                const double pin = 0.5 * dq0[i] * tin;
                travel = std::fabs(pin) + std::fabs(dQ[i] - pin);
            }
        }
        const int s = cs.res[i] != 0 ? (int)(travel / cs.res[i] + 0.99) : (int)(travel * 100);
        total += s;
        if (s > nsteps) { nsteps = s; largest = i; }
    }
    if (total == 0 && start > 0) return count;
    if (start == 0) { push(q0, 0.0, 1); start = 1; }
    if (nsteps == 0) return count;

    const double aL = acc[largest], vL = dq0[largest], fLargestStep = dQ[largest];
    double tInflection = T, fInflection = dQ[largest], fTravel = fInflection, fDelta;
    if (std::fabs(aL) > kEpsilonLinear) {
        const double tin = -vL / aL;
        if (tin >= 0 && tin < T) {
            tInflection = tin;
            fInflection = 0.5 * vL * tInflection;
            fTravel = std::fabs(fInflection) + std::fabs(dQ[largest] - fInflection);
            fDelta = fTravel / nsteps;
            if (fInflection < 0 || (std::fabs(fInflection) <= kEpsilonLinear && dQ[largest] > 0)) fDelta = -fDelta;
        } else fDelta = fTravel / (double)nsteps;
    } else fDelta = fTravel / (double)nsteps;

    double roots[2], timestep = 0, fStep = 0, prevtime = 0, fBest = 0;
    int istep = 0, repeating = 0;
    bool compute_new = true, pending = false;
    while (istep < nsteps && prevtime < T) {
        if (pending) { push(cur.data(), timestep, 1); pending = false; }
        double tmin = prevtime;
        if (compute_new) {
            if (prevtime >= tInflection) fBest = fStep - fDelta;
            else if ((fDelta > 0 && fStep + fDelta > fInflection) || (fDelta < 0 && fStep + fDelta < fInflection)) {
                fBest = fInflection - (fStep + fDelta - fInflection);       // straddles the turning point: reflect
                tmin = tInflection - 1e-7;
            } else fBest = fStep + fDelta;
        }
        if (std::fabs(aL) <= kEpsilonLinear) {
            if (!(std::fabs(vL) > kEpsilon)) return -1;                      // "axis does not move?" (the reference asserts)
            timestep = fBest / vL;
        } else {
            double fNew = fBest;
            bool found = false;
            int nr = solvequad(aL * 0.5, vL, -fNew, roots[0], roots[1]);
            if (nr == 0 && std::fabs(fNew - fLargestStep) < 1e-7) { timestep = T; found = true; }
            if (!found) {
                for (int r = 0; r < nr; ++r)
                    if (roots[r] > tmin && (!found || timestep > roots[r])) { timestep = roots[r]; found = true; }
                if (!found && tmin < tInflection) {
                    fNew = fStep;
                    nr = solvequad(aL * 0.5, vL, -fNew, roots[0], roots[1]);
                    for (int r = 0; r < nr; ++r)
                        if (roots[r] > tmin && (!found || timestep > roots[r])) { timestep = roots[r]; fBest = fNew; found = true; }
                }
            }
            if (!found) return -1;
        }
        prev = cur;
        prevvel = vel;
        double dqscale = 1.0;
        for (int i = 0; i < n; ++i) {
            dQ[i] = q0[i] + timestep * (dq0[i] + timestep * 0.5 * acc[i]) - cur[i];
            if (std::fabs(dQ[i]) > cs.res[i] * 1.01) {
                const int nr = solvequad(acc[i] * 0.5, dq0[i], q0[i] - cur[i] - cs.res[i], roots[0], roots[1]);
                double tscaled = 0;
                bool fs = false;
                for (int r = 0; r < nr; ++r)
                    if (roots[r] >= tmin && roots[r] <= timestep) {
                        if (!fs || roots[r] < tscaled) tscaled = roots[r];
                        fs = true;
                    }
                const double sc = (fs && timestep > prevtime) ? (tscaled - prevtime) / (timestep - prevtime) : std::fabs(cs.res[i] / dQ[i]);
                if (sc < dqscale) dqscale = sc;
            }
        }
        if (dqscale < 1) {
            ++repeating;
            if (dqscale <= 0.01 || repeating > nsteps * 2) return -1;
            timestep = prevtime + (timestep - prevtime) * dqscale;
            for (int i = 0; i < n; ++i) {
                dQ[i] = q0[i] + timestep * (dq0[i] + timestep * 0.5 * acc[i]) - cur[i];
                vel[i] = dq0[i] + timestep * acc[i];
            }
        } else {
            for (int i = 0; i < n; ++i) vel[i] = dq0[i] + timestep * acc[i];
            repeating = 0;
            if (timestep > T + 1e-7) {
                if (istep + 1 >= nsteps) break;                              // expected at the very end
                return -1;
            } else if (timestep > T) timestep = T;
        }
        cs.neigh(cur.data(), dQ.data());                                     // a clamp only sets bHasRampDeviatedFromInterpolation
        pending = true;
        bool moved = false;
        int npost = 1;
        for (int i = 0; i < n; ++i) {
            const double f = std::fabs(cur[i] - prev[i]);
            if (f > cs.res[i] * 1.01) npost = std::max(npost, (int)(f / cs.res[i] + 0.9999));
            if (f > 0.0001) moved = true;
        }
        if (npost > 1) {
            const double inp = 1 / (double)npost;
            std::vector<double> step(n);
            for (int i = 0; i < n; ++i) step[i] = (cur[i] - prev[i]) * inp;
            for (int k = 0; k + 1 < npost; ++k) {
                for (int i = 0; i < n; ++i) prev[i] += step[i];
                push(prev.data(), NAN, 0);
            }
        }
        if (!moved || (istep + 1 < nsteps && repeating > 2) || dqscale >= 1) { fStep = fBest; compute_new = true; ++istep; }
        else compute_new = false;
        prevtime = timestep;
    }
    if (std::fabs(fStep - fLargestStep) > std::fabs(fDelta)) return -1;      // "fStep did not reach fLargestStep"
    {
        int npost = 1;
        for (int i = 0; i < n; ++i) {
            const double f = std::fabs(q1[i] - cur[i]);
            if (f > cs.res[i] * 1.01) npost = std::max(npost, (int)(f / cs.res[i] + 0.9999));
        }
        if (npost > 1) {
            const double inp = 1 / (double)npost;
            std::vector<double> step(n);
            for (int i = 0; i < n; ++i) step[i] = (q1[i] - cur[i]) * inp;
            prev = cur;
            for (int k = 0; k + 1 < npost; ++k) {
                for (int i = 0; i < n; ++i) prev[i] += step[i];
                push(prev.data(), NAN, 2);
            }
        }
    }
    if (pending && cs.dist(cur.data(), q1) > 1e-7) {
        push(cur.data(), timestep, 1);
        // the walk stopped short of q1: the reference then leaves q1 out of the reported configurations (:3280)
        if (check_end && filled) (*filled)[filled->size() - count] = 3;
    }
    return count;
}

// ------------------------------------------------------------------------------------------------ request / coroutine glue
struct Group {                       // one request of one problem: ordered segments of ordered configurations
    std::vector<double> q;
    std::vector<int32_t> seg_begin;  // [nseg + 1], in configurations, local to the group
    bool state_error = false;        // a segment could not be sampled (edge left the limits): treated as a failure of that segment
    int error_segment = -1;
    void clear() { q.clear(); seg_begin.assign(1, 0); state_error = false; error_segment = -1; }
    int nseg() const { return (int)seg_begin.size() - 1; }
};

struct CheckReply { int first_fail_segment = -1; int verdict = 0; };

struct Problem;

template <class T>
struct Task {
    struct promise_type {
        T value{};
        std::coroutine_handle<> continuation;
        Task get_return_object() { return Task{std::coroutine_handle<promise_type>::from_promise(*this)}; }
        std::suspend_always initial_suspend() noexcept { return {}; }
        struct Final {
            bool await_ready() noexcept { return false; }
            std::coroutine_handle<> await_suspend(std::coroutine_handle<promise_type> h) noexcept
            {
                auto c = h.promise().continuation;
                return c ? c : std::noop_coroutine();
            }
            void await_resume() noexcept {}
        };
        Final final_suspend() noexcept { return {}; }
        void return_value(T v) { value = std::move(v); }
        void unhandled_exception() { std::terminate(); }
    };
    std::coroutine_handle<promise_type> h;
    explicit Task(std::coroutine_handle<promise_type> hh) : h(hh) {}
    Task(Task&& o) noexcept : h(o.h) { o.h = nullptr; }
    Task(const Task&) = delete;
    ~Task() { if (h) h.destroy(); }
    bool await_ready() const noexcept { return false; }
    std::coroutine_handle<> await_suspend(std::coroutine_handle<> c) noexcept { h.promise().continuation = c; return h; }
    T await_resume() { return std::move(h.promise().value); }
};

// ------------------------------------------------------------------------------------------------ RRT tree
struct Tree {
    int dof = 0;
    bool fromgoal = false;
    double step = 0.04, mindist = 0;
    std::vector<double> q;          // [nodes][dof]
    std::vector<int32_t> parent;
    std::vector<uint32_t> userdata;
    std::vector<uint8_t> usenn;      // 0: left out of nearest-neighbour queries (InvalidateNodesWithParent, rplanners.h:242-266)
    int size() const { return (int)parent.size(); }
    const double* cfg(int i) const { return &q[(size_t)i * dof]; }
    int nearest(const CSpace& cs, const double* target, double* dist_out = nullptr) const
    {
        int best = -1;
        double bd = INFINITY;
        for (int i = 0; i < size(); ++i) {
            if (!usenn[i]) continue;
            const double d = cs.dist(cfg(i), target);
            if (d < bd) { bd = d; best = i; }
        }
        if (dist_out) *dist_out = bd;
        return best;
    }
    // a goal that has been reached and everything grown from it leave the nearest-neighbour queries (the nodes stay: paths
    // already extracted refer to them, and _InsertNode's minimum-distance rule still sees them)
    void invalidate_with_parent(int root)
    {
        usenn[root] = 0;
        for (int i = root + 1; i < size(); ++i)        // parents precede their children
            if (parent[i] >= 0 && !usenn[parent[i]]) usenn[i] = 0;
    }
    // is some node within `r` of c?  (weighted distance; a node is rejected as soon as its partial sum exceeds r^2, which for
    // r = 0.001 * step happens at the first coordinate for all but coincident nodes)
    bool any_within(const CSpace& cs, const double* c, double r) const
    {
        const double r2 = r * r;
        for (int i = 0; i < size(); ++i) {
            const double* n = cfg(i);
            double s = 0;
            int k = 0;
            for (; k < dof; ++k) {
                double d = n[k] - c[k];
                if (cs.any_circular && cs.circ[k]) {
                    if (d < -kPi) d += 2 * kPi * std::ceil((-kPi - d) / (2 * kPi));
                    else if (d >= kPi) d -= 2 * kPi * std::floor((d + kPi) / (2 * kPi));
                }
                s += cs.w2[k] * d * d;
                if (s > r2) break;
            }
            if (k == dof) return true;
        }
        return false;
    }
    // _InsertNode (rplanners.h:710-738): -1 when a node within _mindistance already exists
    int insert(const CSpace& cs, int par, const double* c, uint32_t ud)
    {
        if (size() > 0 && any_within(cs, c, mindist)) return -1;
        q.insert(q.end(), c, c + dof);
        parent.push_back(par);
        userdata.push_back(ud);
        usenn.push_back(1);
        return size() - 1;
    }
};

enum ExtendType { ET_Failed = 0, ET_Sucess = 1, ET_Connected = 2 };

struct PathList {   // LinearSmoother's list< pair<config, distance-from-previous> >
    struct Node { std::vector<double> q; double d; };
    std::list<Node> nodes;
};

}  // namespace

struct cc_planner {
    CSpace cs;
    cc_query* query = nullptr;
    cc_group_check_fn fn = nullptr;
    void* user = nullptr;
    cc_config_sampler_fn sample_goal_fn = nullptr, sample_initial_fn = nullptr;     // PlannerParameters::_samplegoalfn / _sampleinitialfn
    void* sampler_user = nullptr;
    cc_progress_fn progress_fn = nullptr;                                           // PlannerBase::_CallCallbacks
    void* progress_user = nullptr;
    std::vector<std::unique_ptr<Problem>> problems;
    uint64_t totals[4] = {0, 0, 0, 0};
    // page-locked staging for the GPU checker (cc_host_alloc): configurations in, verdicts out
    double* pin_q = nullptr; size_t pin_q_cap = 0;
    int query_cols = 0;              // columns of the rows the bound query takes: cs.n + its host-evaluated mimic columns
    uint32_t* pin_v = nullptr; size_t pin_v_cap = 0;
    ~cc_planner() { if (pin_q) cc_host_free(pin_q); if (pin_v) cc_host_free(pin_v); }
};

namespace {

struct Problem {
    const CSpace* cs = nullptr;
    cc_birrt_params prm{};
    std::vector<double> start, goal;          // first initial / goal configuration (BasicRRT plans with these)
    std::vector<double> starts_all, goals_all; // every initial / goal configuration of the problem, flat (BiRRT)
    uint32_t seed = 0;
    int index = 0;                            // position in the batch (what the sampler callbacks are told)
    cc_config_sampler_fn sample_goal_fn = nullptr, sample_initial_fn = nullptr;
    void* sampler_user = nullptr;
    cc_progress_fn progress_fn = nullptr;
    void* progress_user = nullptr;
    std::chrono::steady_clock::time_point t_start;
    cc_plan_result res{};
    std::vector<double> path_raw, path;       // flat [waypoints][dof]

    // request / reply mailbox
    Group group;
    CheckReply reply;
    std::coroutine_handle<> waiting;
    bool has_request = false, done = false;

    struct CheckAwait {
        Problem* p;
        bool await_ready() const noexcept { return false; }
        void await_suspend(std::coroutine_handle<> h) noexcept { p->waiting = h; p->has_request = true; }
        CheckReply await_resume() noexcept { return p->reply; }
    };
    // submit the group built in `group`; an empty group needs no checker
    CheckAwait check() { return CheckAwait{this}; }

    Tree fwd, bwd;
    Sampler uniform, config;

    int ndof() const { return cs->n; }

    // one segment = one CheckPathAllConstraints call
    void add_edge(const double* q0, const double* q1, int interval, std::vector<double>* times = nullptr)
    {
        const size_t before = group.q.size();
        const int c = edge_samples(*cs, q0, q1, interval, group.q, times);
        if (c < 0) {
            group.q.resize(before);
            if (!group.state_error) { group.state_error = true; group.error_segment = group.nseg(); }
        }
        group.seg_begin.push_back((int32_t)(group.q.size() / ndof()));
    }

    // RobotConfigurationSampler::SampleSequence (robotconfiguration.cpp:81-112) over its own mt19937
    void sample_config(std::vector<double>& out)
    {
        out.resize(ndof());
        for (int i = 0; i < ndof(); ++i) {
            const double u = config.real(IT_Closed);
            out[i] = cs->circ[i] ? -kPi + 2 * kPi * u : cs->lo[i] + u * (cs->hi[i] - cs->lo[i]);
        }
    }

    // SpatialTree::Extend.  The ray c0 -> c1 -> ... is generated up front (it depends on the target only); chunks of
    // `speculate_steps` steps are submitted, one segment per step, and nodes are inserted for the steps before the first
    // failing one.
    Task<int> extend(Tree& tree, const std::vector<double>& target, int& lastnode, bool one_step = false)
    {
        const int n = ndof();
        const int nn = tree.nearest(*cs, target.data());
        if (nn < 0) co_return (int)ET_Failed;
        int pnode = nn;
        lastnode = nn;
        bool added = false;
        std::vector<double> cur(tree.cfg(nn), tree.cfg(nn) + n), nxt(n), delta(n);
        const int maxsteps = prm.max_extend_steps > 0 ? prm.max_extend_steps : 100;
        // steps submitted per round: the whole ray (0), a fixed number (> 0), or adaptive (< 0): -speculate_steps steps first,
        // doubling every time a chunk passes — rays that die at an obstacle cost little, free rays finish in a few rounds
        int chunk = prm.speculate_steps > 0 ? prm.speculate_steps : (prm.speculate_steps < 0 ? -prm.speculate_steps : maxsteps);
        if (one_step) chunk = 1;          // bOneStep (rplanners.h:442-444): the ray ends at its first inserted node
        int iter = 0;
        int terminal = -1;                      // set when the ray ends without needing a check
        bool first_chunk = true;
        while (iter < maxsteps && terminal < 0) {
            if (!first_chunk && prm.speculate_steps < 0 && !one_step) chunk *= 2;
            first_chunk = false;
            // ---- generate up to `chunk` steps
            group.clear();
            std::vector<double> ray;            // the new configuration of every generated step
            std::vector<double> c = cur;
            int end_code = -1;                  // how the generation stopped: -1 = chunk/bound exhausted
            int gen = 0;
            while (gen < chunk && iter + gen < maxsteps) {
                double fdist = cs->dist(c.data(), target.data());
                if (fdist > tree.step) fdist = tree.step / fdist;
                else if (fdist <= 0.01 * tree.step) { end_code = ET_Connected; break; }
                else fdist = 1;
                nxt = c;
                delta = target;
                cs->diff(delta.data(), c.data());
                for (int i = 0; i < n; ++i) delta[i] *= fdist;
                cs->neigh(nxt.data(), delta.data());            // NSS_Failed never happens for the default neighbour function
                if (cs->dist(c.data(), nxt.data()) <= 0.01 * tree.step) { end_code = 100; break; }   // did not move
                if (tree.fromgoal) add_edge(nxt.data(), c.data(), IT_OpenEnd);
                else add_edge(c.data(), nxt.data(), IT_OpenStart);
                ray.insert(ray.end(), nxt.begin(), nxt.end());
                c = nxt;
                ++gen;
            }
            int passed = gen;
            if (gen > 0) {
                CheckReply r = co_await check();
                if (r.first_fail_segment >= 0) passed = r.first_fail_segment;
            }
            for (int s = 0; s < passed; ++s) {
                const int id = tree.insert(*cs, pnode, &ray[(size_t)s * n], 0);
                if (id >= 0) { pnode = id; lastnode = id; added = true; }
            }
            if (passed < gen) { terminal = added ? ET_Sucess : ET_Failed; break; }
            if (one_step && added) { terminal = ET_Connected; break; }    // "BasicRRT relies on ET_Connected" (rplanners.h:443)
            iter += gen;
            cur = c;
            if (end_code == ET_Connected) { terminal = ET_Connected; break; }
            if (end_code == 100) { terminal = added ? ET_Sucess : ET_Failed; break; }
        }
        if (terminal < 0) terminal = added ? ET_Sucess : ET_Failed;
        co_return terminal;
    }

    double path_length(const std::vector<double>& p) const
    {
        const int n = ndof();
        double L = 0;
        for (size_t i = 1; i < p.size() / n; ++i) L += cs->dist(&p[(i - 1) * n], &p[i * n]);
        return L;
    }

    // RrtPlanner::_SimpleOptimizePath (rrt.h:151-240) on a flat path
    Task<int> simple_optimize(std::vector<double>& p, int numiterations)
    {
        const int n = ndof();
        if ((int)p.size() <= 2 * n) co_return 0;
        int nrejected = 0, curiter = numiterations;
        std::vector<double> vstart(n), vend(n), d0(n), d1(n), times;
        while (curiter > 0 && nrejected < (int)p.size() + 4 * n) {
            --curiter;
            const int npts = (int)p.size() / n;
            const int endIndex = 2 + (int)(uniform.u32() % (uint32_t)(npts - 2));
            const int startIndex = (int)(uniform.u32() % (uint32_t)(endIndex - 1));
            nrejected++;
            std::copy(p.begin() + (size_t)startIndex * n, p.begin() + (size_t)(startIndex + 1) * n, vstart.begin());
            std::copy(p.begin() + (size_t)(endIndex - 1) * n, p.begin() + (size_t)endIndex * n, vend.begin());
            d0 = vend;
            cs->diff(d0.data(), vstart.data());
            const int mid = startIndex + (endIndex - startIndex) / 2;
            std::copy(p.begin() + (size_t)mid * n, p.begin() + (size_t)(mid + 1) * n, d1.begin());
            cs->diff(d1.data(), vstart.data());
            double dot = 0, l0 = 0, l1 = 0;
            for (int i = 0; i < n; ++i) { dot += d0[i] * d1[i]; l0 += d0[i] * d0[i]; l1 += d1[i] * d1[i]; }
            if (!(std::fabs(dot * dot - l0 * l1) > kEpsilonDotProduct)) continue;       // colinear: nothing to gain
            group.clear();
            times.clear();
            add_edge(vstart.data(), vend.data(), IT_Open, &times);
            bool failed = group.state_error;
            if (!failed && group.q.size() > 0) {
                CheckReply r = co_await check();
                failed = r.first_fail_segment >= 0;
            }
            if (failed) {
                if (nrejected++ > (int)p.size() + 8) break;
                continue;
            }
            // the checked configurations replace [startIndex+1, endIndex) — the slot of vend included (rrt.h:214-231)
            std::vector<double> np(p.begin(), p.begin() + (size_t)(startIndex + 1) * n);
            np.insert(np.end(), group.q.begin(), group.q.end());
            np.insert(np.end(), p.begin() + (size_t)endIndex * n, p.end());
            p.swap(np);
            nrejected = 0;
            if ((int)p.size() <= 2 * n) co_return 0;
        }
        co_return 0;
    }

    // LinearSmoother::SegmentFeasible (linearsmoother.cpp:1235-1263): the segment and its +/- pointtolerance * resolution copies
    void add_feasible_segments(const double* a, const double* b, int interval)
    {
        const int n = ndof();
        std::vector<double> an(n), bn(n);
        const double pert[3] = {0, prm.point_tolerance, -prm.point_tolerance};
        for (double f : pert) {
            for (int i = 0; i < n; ++i) {
                an[i] = std::min(std::max(a[i] + f * cs->res[i], cs->lo[i]), cs->hi[i]);
                bn[i] = std::min(std::max(b[i] + f * cs->res[i], cs->lo[i]), cs->hi[i]);
            }
            add_edge(an.data(), bn.data(), interval);
        }
    }

    struct Span {   // the random sub-path both smoother passes pick (linearsmoother.cpp:293-325)
        std::list<PathList::Node>::iterator itstart, itstartprev, itend, itendprev;
        double fstartdist, fenddist, fstartdelta, fenddelta;
        int numnodes;
        bool degenerate;
    };
    Span pick_span(PathList& lp, double totaldist, Sampler& rs) const
    {
        Span s;
        s.fstartdist = std::max(0.0, totaldist - prm.step_length) * rs.real(IT_OpenEnd);
        s.fenddist = s.fstartdist + (totaldist - s.fstartdist) * rs.real(IT_OpenStart);
        s.fstartdelta = s.fenddelta = 0;
        double fcur = 0;
        s.itstartprev = s.itstart = lp.nodes.begin();
        while (s.itstart != lp.nodes.end()) {
            if (s.fstartdist >= fcur && s.fstartdist < fcur + s.itstart->d) { s.fstartdelta = s.fstartdist - fcur; break; }
            fcur += s.itstart->d;
            s.itstartprev = s.itstart;
            ++s.itstart;
        }
        s.itendprev = s.itstartprev;
        s.itend = s.itstart;
        s.numnodes = 0;
        while (s.itend != lp.nodes.end()) {
            if (s.fenddist >= fcur && s.fenddist < fcur + s.itend->d) { s.fenddelta = s.fenddist - fcur; break; }
            fcur += s.itend->d;
            s.itendprev = s.itend;
            ++s.itend;
            ++s.numnodes;
        }
        s.degenerate = (s.itstart == s.itend) || s.itstart == lp.nodes.end() || s.itend == lp.nodes.end();
        return s;
    }

    // LinearSmoother::_OptimizePath (linearsmoother.cpp:279-383)
    Task<double> smooth_shortcut(PathList& lp, double totaldist, int maxiter, Sampler& rs)
    {
        const int n = ndof();
        std::vector<double> vs(n), ve(n);
        int nrejected = 0;
        for (int it = 0; it < maxiter; ++it) {
            if (nrejected >= 20) break;
            Span s = pick_span(lp, totaldist, rs);
            if (s.degenerate) continue;
            nrejected++;
            if (s.itstart == lp.nodes.begin()) vs = s.itstart->q;
            else {
                const double f = s.fstartdelta / s.itstart->d;
                for (int i = 0; i < n; ++i) vs[i] = s.itstart->q[i] * f + s.itstartprev->q[i] * (1 - f);
            }
            if (s.itend == std::prev(lp.nodes.end())) ve = s.itend->q;
            else {
                const double f = s.fenddelta / s.itend->d;
                for (int i = 0; i < n; ++i) ve[i] = s.itend->q[i] * f + s.itendprev->q[i] * (1 - f);
            }
            const double newdist = cs->dist(vs.data(), ve.data());
            if (newdist > s.fenddist - s.fstartdist - 0.5 * prm.step_length) continue;
            group.clear();
            add_feasible_segments(vs.data(), ve.data(), IT_Open);
            bool failed = group.state_error;
            if (!failed && group.q.size() > 0) {
                CheckReply r = co_await check();
                failed = r.first_fail_segment >= 0;
            }
            if (failed) continue;
            lp.nodes.insert(s.itstart, PathList::Node{vs, s.fstartdelta});
            s.itend->d -= s.fenddelta;
            s.itend = lp.nodes.insert(s.itend, PathList::Node{ve, newdist});
            lp.nodes.erase(s.itstart, s.itend);
            totaldist += newdist - (s.fenddist - s.fstartdist);
            nrejected = 0;
        }
        co_return totaldist;
    }

    // LinearSmoother::_OptimizePathSingleDOF (linearsmoother.cpp:385-559): straighten ONE DOF over a random sub-path
    Task<double> smooth_single_dof(PathList& lp, double totaldist, int maxiter, Sampler& rs)
    {
        const int n = ndof();
        std::vector<PathList::Node> pv;
        int nrejected = 0;
        for (int it = 0; it < maxiter; ++it) {
            if (nrejected >= 20) break;
            Span s = pick_span(lp, totaldist, rs);
            const uint32_t idof = rs.u32() % (uint32_t)n;
            if (s.degenerate) continue;
            nrejected++;
            pv.assign(s.numnodes + 2, PathList::Node{std::vector<double>(n), 0});
            if (std::fabs(s.fstartdelta) <= kEpsilonLinear) pv[0] = PathList::Node{s.itstart->q, 0};
            else {
                const double f = s.fstartdelta / s.itstart->d;
                for (int i = 0; i < n; ++i) pv[0].q[i] = s.itstart->q[i] * f + s.itstartprev->q[i] * (1 - f);
            }
            if (std::fabs(s.fenddelta - s.itend->d) <= kEpsilonLinear) pv[s.numnodes + 1] = *s.itend;
            else {
                pv[s.numnodes + 1].d = s.fenddelta;
                const double f = s.fenddelta / s.itend->d;
                for (int i = 0; i < n; ++i) pv[s.numnodes + 1].q[i] = s.itend->q[i] * f + s.itendprev->q[i] * (1 - f);
            }
            const double v0 = pv[0].q[idof], v1 = pv[s.numnodes + 1].q[idof];
            const double fdelta = (v1 - v0) / (s.fenddist - s.fstartdist);
            double fcur = 0, newdist = 0;
            auto itn = s.itstart;
            int k = 0;
            do {
                if (k == 0) fcur = itn->d - s.fstartdelta; else fcur += itn->d;
                pv[k + 1].q = itn->q;
                pv[k + 1].q[idof] = v0 + fcur * fdelta;
                pv[k + 1].d = cs->dist(pv[k].q.data(), pv[k + 1].q.data());
                newdist += pv[k + 1].d;
                ++itn;
                ++k;
            } while (itn != s.itend);
            pv[s.numnodes + 1].d = cs->dist(pv[s.numnodes].q.data(), pv[s.numnodes + 1].q.data());
            newdist += pv[s.numnodes + 1].d;
            if (newdist > s.fenddist - s.fstartdist - 0.5 * prm.step_length) continue;
            group.clear();
            for (size_t i = 0; i + 1 < pv.size(); ++i)
                add_feasible_segments(pv[i].q.data(), pv[i + 1].q.data(), i + 2 == pv.size() ? IT_Open : IT_OpenStart);
            bool failed = group.state_error;
            if (!failed && group.q.size() > 0) {
                CheckReply r = co_await check();
                failed = r.first_fail_segment >= 0;
            }
            if (failed) continue;
            if (std::fabs(s.fstartdelta) > kEpsilonLinear) {
                pv[0].d = s.fstartdelta;
                lp.nodes.insert(s.itstart, pv[0]);
            }
            itn = s.itstart;
            k = 1;
            do { *itn = pv[k++]; ++itn; } while (itn != s.itend);
            if (std::fabs(s.fenddelta - s.itend->d) > kEpsilonLinear) {
                lp.nodes.insert(s.itend, pv[s.numnodes + 1]);
                s.itend->d -= s.fenddelta;
            }
            totaldist += newdist - (s.fenddist - s.fstartdist);
            nrejected = 0;
        }
        co_return totaldist;
    }

    // LinearSmoother::PlanPath (linearsmoother.cpp:104-262), _nUseSingleDOFSmoothing == 1: 8/10 shortcuts, 2/10 single-DOF
    Task<int> smooth(const std::vector<double>& in, std::vector<double>& out)
    {
        const int n = ndof();
        Sampler rs(seed);
        PathList lp;
        double total = 0;
        lp.nodes.push_back(PathList::Node{std::vector<double>(in.begin(), in.begin() + n), 0});
        for (size_t i = 1; i < in.size() / n; ++i) {
            const double d = cs->dist(lp.nodes.back().q.data(), &in[i * n]);
            if (d > 0) { lp.nodes.push_back(PathList::Node{std::vector<double>(in.begin() + i * n, in.begin() + (i + 1) * n), d}); total += d; }
        }
        const int iters = prm.smoother_iterations;
        if (lp.nodes.size() > 1) {
            if (prm.post_processor == CC_POST_LINEARSMOOTHER_SHORTCUTS) {
                co_await smooth_shortcut(lp, total, iters, rs);               // "linearsmoother 0": every iteration is a shortcut (:241-244)
            } else {
                const double d1 = co_await smooth_shortcut(lp, total, iters * 8 / 10, rs);
                co_await smooth_single_dof(lp, d1, iters * 2 / 10, rs);
            }
        }
        out.clear();
        for (auto& nd : lp.nodes) out.insert(out.end(), nd.q.begin(), nd.q.end());
        if (lp.nodes.size() <= 1) out.insert(out.end(), in.end() - n, in.end());
        co_return 0;
    }

    // ---------------------------------------------------------------------------------------- shortcut_linear post-processor
    // ShortcutLinearPlanner (plugins/rplanners/linearshortcutadvanced.cpp): the planner basemanipulation selects
    // (rmanipulation/basemanipulation.cpp:788).  Every CheckPathAllConstraints call carries CFO_FillCheckedConfiguration, and
    // the configurations it returns on success are the samples it visited (planningutils.cpp:2801-2813, 3410-3413, 3524-3533,
    // 3611-3623): q0 first when the interval is closed at the start, the interior samples, q1 LAST when it is closed at the end
    // — i.e. edge_samples' checking order with the leading q1 rotated to the back.  They do not depend on any verdict, so a
    // whole shortcut (all segments of a single-DOF shortcut) is generated first and sent to the checker as one group.
    void filled_configurations(size_t cfg_begin, size_t cfg_end, int interval, std::vector<double>& out) const
    {
        const int n = ndof();
        out.clear();
        if (cfg_end <= cfg_begin) return;
        const bool check_end = (interval == IT_OpenStart || interval == IT_Closed);
        const double* q = group.q.data();
        if (check_end) {
            out.insert(out.end(), q + (cfg_begin + 1) * n, q + cfg_end * n);
            out.insert(out.end(), q + cfg_begin * n, q + (cfg_begin + 1) * n);
        } else out.insert(out.end(), q + cfg_begin * n, q + cfg_end * n);
    }

    // ShortcutLinearPlanner::_SubsampleTrajectory (linearshortcutadvanced.cpp:578-652)
    void subsample_trajectory(const std::vector<double>& in, PathList& lp) const
    {
        const int n = ndof();
        const size_t nw = in.size() / n;
        std::vector<double> q0(in.begin(), in.begin() + n), q1(n), dq(n), dq2(n), qcur;
        lp.nodes.push_back(PathList::Node{q0, 0});
        qcur = q0;
        for (size_t ip = 1; ip < nw; ++ip) {
            q1.assign(in.begin() + ip * n, in.begin() + (ip + 1) * n);
            dq = q1;
            cs->diff(dq.data(), q0.data());
            int nsteps = 1;
            for (int i = 0; i < n; ++i) {
                const int s = cs->res[i] != 0 ? (int)(std::fabs(dq[i]) / cs->res[i]) : (int)(std::fabs(dq[i]) * 100);
                nsteps = std::max(nsteps, s);
            }
            const double fisteps = 1.0 / nsteps;
            for (int i = 0; i < n; ++i) dq[i] *= fisteps;
            // q + dq repeatedly; the default neighbour function never fails, so the `mult` retry of the reference stays at 1;
            // a clamped step (NSS_SuccessfulWithDeviation) ends the subsampling of this segment
            for (int f = 1; f < nsteps; ++f) {
                if (cs->neigh(qcur.data(), dq.data())) { qcur = lp.nodes.back().q; break; }
                const double d = cs->dist(lp.nodes.back().q.data(), qcur.data());
                lp.nodes.push_back(PathList::Node{qcur, d});
            }
            const double d = cs->dist(lp.nodes.back().q.data(), q1.data());
            lp.nodes.push_back(PathList::Node{q1, d});
            qcur = q1;
            q0.swap(q1);
        }
        // the reference appends the last waypoint once more (:648-650), at distance 0 from its twin
        q1.assign(in.end() - n, in.end());
        lp.nodes.push_back(PathList::Node{q1, cs->dist(lp.nodes.back().q.data(), q1.data())});
    }

    // ShortcutLinearPlanner::_OptimizePath (linearshortcutadvanced.cpp:188-339)
    Task<int> shortcut_linear_path(PathList& lp, int maxiter, Sampler& rs)
    {
        const int n = ndof();
        int numshortcuts = 0;
        std::vector<double> cfgs, dists;
        for (int iiter = maxiter; iiter > 0 && lp.nodes.size() > 2;) {
            --iiter;
            const uint32_t endIndex = 2 + rs.u32() % ((uint32_t)lp.nodes.size() - 2);
            const uint32_t startIndex = rs.u32() % (endIndex - 1);
            auto itstart = std::next(lp.nodes.begin(), startIndex);
            auto itend = itstart;
            double total = 0;
            for (uint32_t j = 0; j < endIndex - startIndex; ++j) { ++itend; total += itend->d; }
            if (cs->dist(itstart->q.data(), itend->q.data()) > total - 0.1 * prm.step_length) continue;
            group.clear();
            add_edge(itstart->q.data(), itend->q.data(), IT_Open);
            if (group.state_error) continue;
            filled_configurations(0, group.q.size() / n, IT_Open, cfgs);
            if (cfgs.empty()) continue;                        // nothing between the end points: the reference rejects (:265-270)
            CheckReply r = co_await check();
            if (r.first_fail_segment >= 0) continue;
            const size_t nc = cfgs.size() / n;
            dists.assign(nc + 1, 0.0);
            double newtotal = dists[0] = cs->dist(itstart->q.data(), &cfgs[0]);
            for (size_t k = 1; k < nc; ++k) { dists[k] = cs->dist(&cfgs[(k - 1) * n], &cfgs[k * n]); newtotal += dists[k]; }
            dists[nc] = cs->dist(&cfgs[(nc - 1) * n], itend->q.data());
            newtotal += dists[nc];
            if (newtotal > total - 0.1 * prm.step_length) continue;
            ++itstart;
            for (size_t k = 0; k < nc; ++k)
                lp.nodes.insert(itstart, PathList::Node{std::vector<double>(cfgs.begin() + k * n, cfgs.begin() + (k + 1) * n), dists[k]});
            itend->d = dists[nc];
            lp.nodes.erase(itstart, itend);
            ++numshortcuts;
        }
        co_return numshortcuts;
    }

    // ShortcutLinearPlanner::_OptimizePathOneDOF (linearshortcutadvanced.cpp:344-569): one DOF is interpolated linearly in
    // path length over a random sub-path, the other DOFs keep their values
    Task<int> shortcut_linear_one_dof(PathList& lp, int maxiter, Sampler& rs)
    {
        const int n = ndof();
        int numshortcuts = 0;
        std::vector<double> vcur(n), vnext(n), vdelta(n), cfgs;
        std::list<PathList::Node> sp;
        for (int iiter = 0; iiter < maxiter; ++iiter) {
            if (lp.nodes.size() <= 2) break;
            uint32_t idof = rs.u32() % (uint32_t)(n + 1);       // the last DOF gets twice the chance (:374-377)
            if (idof > (uint32_t)(n - 1)) idof = n - 1;
            const uint32_t endIndex = 2 + rs.u32() % ((uint32_t)lp.nodes.size() - 2);
            const uint32_t startIndex = rs.u32() % (endIndex - 1);
            if (endIndex == startIndex + 1) continue;
            auto itstart = std::next(lp.nodes.begin(), startIndex);
            auto itend = itstart;
            double totaldof = 0, total = 0, curv = itstart->q[idof];
            for (uint32_t j = 0; j < endIndex - startIndex; ++j) {
                ++itend;
                totaldof += std::fabs(itend->q[idof] - curv);
                total += itend->d;
                curv = itend->q[idof];
            }
            const double expected = itend->q[idof] - itstart->q[idof];
            if (std::fabs(expected) > totaldof - 0.1 * cs->res[idof]) continue;
            sp.clear();
            group.clear();
            bool ok = true;
            const double fdelta = expected / total, fstart = itstart->q[idof];
            double fcurdist = 0, newdof = 0, newtotal = 0;
            auto itnode = itstart;
            vcur = itnode->q;
            do {
                ++itnode;
                fcurdist += itnode->d;
                vnext = itnode->q;
                if (itnode != itend) vnext[idof] = fstart + fcurdist * fdelta;
                // through _neighstatefn from the last accepted configuration (:436-474)
                const std::vector<double>& from = sp.empty() ? vcur : sp.back().q;
                for (int k = 0; k < n; ++k) { vdelta[k] = vnext[k] - from[k]; vnext[k] = from[k]; }
                const bool deviated = cs->neigh(vnext.data(), vdelta.data());
                if (itnode == itend && deviated) { ok = false; break; }
                const size_t c0 = group.q.size() / n;
                add_edge(vcur.data(), vnext.data(), IT_OpenStart);
                if (group.state_error) { ok = false; break; }
                filled_configurations(c0, group.q.size() / n, IT_OpenStart, cfgs);
                if (cfgs.empty()) { ok = false; break; }
                // NB the reference measures the first configuration of EVERY segment from the shortcut's start node (:506-508)
                sp.push_back(PathList::Node{std::vector<double>(cfgs.begin(), cfgs.begin() + n), cs->dist(itstart->q.data(), &cfgs[0])});
                newdof += std::fabs(cfgs[idof] - itstart->q[idof]);
                newtotal += sp.back().d;
                for (size_t k = 1; k < cfgs.size() / n && ok; ++k) {
                    const double* c = &cfgs[k * n];
                    newdof += std::fabs(c[idof] - sp.back().q[idof]);
                    if (newdof > totaldof - 0.1 * cs->res[idof]) { ok = false; break; }
                    const double d = cs->dist(sp.back().q.data(), c);
                    sp.push_back(PathList::Node{std::vector<double>(c, c + n), d});
                    newtotal += d;
                    if (newtotal > 1.1 * total) { ok = false; break; }
                }
                if (!ok) break;
                vcur.swap(vnext);
            } while (itnode != itend);
            if (!ok || sp.empty()) continue;                    // rejected on lengths alone: no verdict can change that
            CheckReply r = co_await check();
            if (r.first_fail_segment >= 0) continue;
            ++numshortcuts;
            ++itstart;
            ++itend;
            lp.nodes.insert(itstart, sp.begin(), sp.end());
            lp.nodes.erase(itstart, itend);
        }
        co_return numshortcuts;
    }

    // ShortcutLinearPlanner::PlanPath (linearshortcutadvanced.cpp:76-184) without the retimer that follows it
    Task<int> shortcut_linear(const std::vector<double>& in, std::vector<double>& out)
    {
        const int n = ndof();
        if ((int)in.size() < 2 * n) { out = in; co_return 0; }
        Sampler rs(seed);                                       // _puniformsampler->SetSeed(_nRandomGeneratorSeed) (:62)
        PathList lp;
        subsample_trajectory(in, lp);
        const int iters = prm.smoother_iterations > 0 ? prm.smoother_iterations : 100;   // _InitPlan (:54-56)
        res.shortcuts = co_await shortcut_linear_path(lp, iters, rs);
        res.shortcuts += co_await shortcut_linear_one_dof(lp, iters, rs);
        out.clear();
        for (auto& nd : lp.nodes) out.insert(out.end(), nd.q.begin(), nd.q.end());
        co_return 0;
    }

    // BasicRrtPlanner::InitPlan + PlanPath (plugins/rplanners/rrt.h:693-938) for one start and one goal: a single forward tree,
    // the goal sampled at the first iteration and then with probability goal_bias (0.05 there), ET_Connected within
    // 2 x step of the goal ends the search (_nMinIterations = 0), the path is the chain from the root to that node (it ends
    // at the node, not at the goal, like the reference's), then _SimpleOptimizePath(10) and the post-processor.  The
    // reference draws from RaveRandomFloat / RaveRandomInt here (global generator): the stream is this problem's sampler.
    Task<int> plan_basic(bool one_step)
    {
        const int n = ndof();
        uniform = Sampler(seed);
        config = Sampler(seed);
        fwd.dof = n;
        fwd.step = prm.step_length;
        fwd.mindist = 0.001 * prm.step_length;
        fwd.fromgoal = false;
        group.clear();
        add_edge(start.data(), start.data(), IT_OpenStart);
        CheckReply r = co_await check();
        if (r.first_fail_segment >= 0) { res.status = CC_PLAN_BAD_INITIAL; co_return 0; }
        fwd.insert(*cs, -1, start.data(), 0);
        group.clear();
        add_edge(goal.data(), goal.data(), IT_OpenStart);                     // goals in collision are dropped (:738-743)
        r = co_await check();
        if (r.first_fail_segment >= 0) { res.status = CC_PLAN_BAD_GOAL; co_return 0; }
        const int maxiter = prm.max_iterations > 0 ? prm.max_iterations : 10000;
        int iter = 0, best = -1, last = -1;
        std::vector<double> target;
        while (iter < maxiter) {
            ++iter;
            res.iterations++;
            if (best >= 0) break;
            if (iter == 1 || uniform.real(IT_Closed) < prm.goal_bias) {
                (void)uniform.u32();                                          // RaveRandomInt() % _vecGoals.size(), one goal
                target = goal;
            } else sample_config(target);
            const int et = co_await extend(fwd, target, last, one_step);
            if (et == ET_Connected && cs->dist(goal.data(), fwd.cfg(last)) < 2 * prm.step_length) best = last;
        }
        res.nodes_forward = fwd.size();
        res.nodes_backward = 0;
        if (best < 0) { res.status = CC_PLAN_FAILED; co_return 0; }
        std::vector<int> chain;
        for (int i = best; i >= 0; i = fwd.parent[i]) chain.push_back(i);
        path_raw.clear();
        for (auto it = chain.rbegin(); it != chain.rend(); ++it) path_raw.insert(path_raw.end(), fwd.cfg(*it), fwd.cfg(*it) + n);
        co_await simple_optimize(path_raw, prm.path_optimize_iterations);
        res.waypoints_raw = (int)path_raw.size() / n;
        res.length_raw = path_length(path_raw);
        if (prm.smoother_iterations > 0 && prm.post_processor == CC_POST_SHORTCUT_LINEAR) co_await shortcut_linear(path_raw, path);
        else if (prm.smoother_iterations > 0) co_await smooth(path_raw, path);
        else path = path_raw;
        res.waypoints = (int)path.size() / n;
        res.length = path_length(path);
        res.status = CC_PLAN_SOLVED;
        co_return 0;
    }

    // BirrtPlanner::InitPlan + PlanPath for one start and one goal
    Task<int> plan()
    {
        const int n = ndof();
        uniform = Sampler(seed);
        config = Sampler(seed);
        for (Tree* t : {&fwd, &bwd}) {
            t->dof = n;
            t->step = prm.step_length;
            t->mindist = 0.001 * prm.step_length;
        }
        fwd.fromgoal = false;
        bwd.fromgoal = true;
        // initial and goal configurations: CheckPathAllConstraints(v, v, IT_OpenStart) each (rrt.h:76-88, 327-343).  Valid initial
        // configurations become roots of the forward tree (userdata = their rank among the valid ones), goals roots of the
        // backward tree (userdata = their index in vgoalconfig; a goal that fails keeps its slot as a dummy)
        CheckReply r;
        const int nstarts = (int)starts_all.size() / n, ngoals = (int)goals_all.size() / n;
        for (int k = 0; k < nstarts; ++k) {
            const double* v = &starts_all[(size_t)k * n];
            group.clear();
            add_edge(v, v, IT_OpenStart);
            r = co_await check();
            if (r.first_fail_segment >= 0) continue;
            fwd.insert(*cs, -1, v, (uint32_t)fwd.size());
        }
        if (fwd.size() == 0) { res.status = CC_PLAN_BAD_INITIAL; co_return 0; }
        std::vector<int> goal_nodes(ngoals, -1);
        int nvalidgoals = 0;
        for (int k = 0; k < ngoals; ++k) {
            const double* v = &goals_all[(size_t)k * n];
            group.clear();
            add_edge(v, v, IT_OpenStart);
            r = co_await check();
            if (r.first_fail_segment >= 0) continue;
            goal_nodes[k] = bwd.insert(*cs, -1, v, (uint32_t)k);
            if (goal_nodes[k] >= 0) ++nvalidgoals;
        }
        if (bwd.size() == 0 && !sample_goal_fn) { res.status = CC_PLAN_BAD_GOAL; co_return 0; }     // "no goals specified" (rrt.h:349-355)

        const int maxiter = prm.max_iterations > 0 ? prm.max_iterations : 10000;
        Tree* A = &fwd;
        Tree* B = &bwd;
        int ia = -1, ib = -1, iter = 0;
        // RRTParameters::_minimumgoalpaths (plannerparameters.h:740): keep going until that many goals — or every valid one — have a
        // path; the shortest is returned (rrt.h:495-531, 540-546)
        const int want_paths = std::max(1, (int)prm.minimum_goal_paths);
        struct GoalPath { std::vector<double> q; double length; int32_t start_index, goal_index; };
        std::vector<GoalPath> goal_paths;
        bool sample_goal = true, interrupted = false;
        int ninitial = fwd.size();                          // _vecInitialNodes.size(): the userdata of the next sampled initial
        std::vector<double> target;
        while (iter < 3 * maxiter) {
            ++iter;
            res.iterations++;
            // callbacks first, "since code can continue" (rrt.h:401-412): PA_Interrupt ends the problem, PA_ReturnWithAnySolution
            // ends the search as soon as some goal has a path; then the time budget (_nMaxPlanningTime, rrt.h:414-420)
            if (progress_fn) {
                const int action = progress_fn(progress_user, index, iter / 3);
                if (action == CC_PA_INTERRUPT) { interrupted = true; break; }
                if (action == CC_PA_RETURN_WITH_ANY_SOLUTION && !goal_paths.empty()) break;
            }
            if (prm.max_planning_time_ms > 0 &&
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count() >= (double)prm.max_planning_time_ms) break;
            // goals and initial configurations that arrive while planning (an IK sampler that keeps producing solutions): added
            // as new roots as they are, like the reference does (rrt.h:422-437); a configuration too close to an existing node
            // keeps its slot as a dummy
            if (sample_goal_fn) {
                std::vector<double> v(n);
                if (sample_goal_fn(sampler_user, index, n, v.data())) {
                    goal_nodes.push_back(bwd.insert(*cs, -1, v.data(), (uint32_t)goal_nodes.size()));
                    ++nvalidgoals;
                }
            }
            if (sample_initial_fn) {
                std::vector<double> v(n);
                if (sample_initial_fn(sampler_user, index, n, v.data())) fwd.insert(*cs, -1, v.data(), (uint32_t)ninitial++);
            }
            if (bwd.size() == 0) continue;                  // nothing to grow towards yet
            const int ngoals_now = (int)goal_nodes.size();
            target.clear();
            if ((sample_goal || uniform.real(IT_Closed) < prm.goal_bias) && nvalidgoals > 0) {
                sample_goal = false;
                // up to 3 x #goals draws until one names a goal that stands and has no path yet (rrt.h:443-466)
                for (int testiter = 0; testiter < ngoals_now * 3; ++testiter) {
                    const uint32_t gi = uniform.u32() % (uint32_t)ngoals_now;
                    if (goal_nodes[gi] < 0) continue;
                    bool found = false;
                    for (const GoalPath& gp : goal_paths) if (gp.goal_index == (int32_t)gi) { found = true; break; }
                    if (found) continue;
                    target.assign(bwd.cfg(goal_nodes[gi]), bwd.cfg(goal_nodes[gi]) + n);
                    break;
                }
            }
            if (target.empty()) sample_config(target);
            int et = co_await extend(*A, target, ia);
            if (et == ET_Failed) {
                if (iter > 3 * maxiter) break;
                continue;
            }
            std::vector<double> toward(A->cfg(ia), A->cfg(ia) + n);
            et = co_await extend(*B, toward, ib);
            if (et == ET_Connected) {
                // _ExtractPath: forward chain root -> connection, then backward chain connection -> root, _SimpleOptimizePath(10)
                const int cf = (A == &fwd) ? ia : ib, cb = (A == &bwd) ? ia : ib;
                GoalPath gp;
                gp.goal_index = -1;
                std::vector<int> chain;
                for (int i = cf; i >= 0; i = fwd.parent[i]) chain.push_back(i);
                for (auto it = chain.rbegin(); it != chain.rend(); ++it) gp.q.insert(gp.q.end(), fwd.cfg(*it), fwd.cfg(*it) + n);
                int goal_root = -1;
                for (int i = cb; i >= 0; i = bwd.parent[i]) {
                    gp.q.insert(gp.q.end(), bwd.cfg(i), bwd.cfg(i) + n);
                    if (bwd.parent[i] < 0) { gp.goal_index = (int32_t)bwd.userdata[i]; goal_root = i; }    // GOALPATH::goalindex (rrt.h:633-644)
                }
                gp.start_index = (int32_t)fwd.userdata[chain.back()];                                      // GOALPATH::startindex (:622-631)
                co_await simple_optimize(gp.q, prm.path_optimize_iterations);
                gp.length = path_length(gp.q);
                goal_paths.push_back(std::move(gp));
                if ((int)goal_paths.size() >= want_paths || (int)goal_paths.size() >= nvalidgoals) break;
                sample_goal = true;
                // more goals wanted: nothing may connect to the tree grown from this goal again
                bwd.invalidate_with_parent(goal_root);
            }
            std::swap(A, B);
            iter += 3;
            if (iter > 3 * maxiter) break;
        }
        res.nodes_forward = fwd.size();
        res.nodes_backward = bwd.size();
        res.goal_paths = (int32_t)goal_paths.size();
        if (interrupted) { res.status = CC_PLAN_INTERRUPTED; co_return 0; }        // PS_Interrupted: no trajectory, whatever was found
        if (goal_paths.empty()) { res.status = CC_PLAN_FAILED; co_return 0; }
        size_t ibest = 0;
        for (size_t k = 1; k < goal_paths.size(); ++k) if (goal_paths[k].length < goal_paths[ibest].length) ibest = k;
        path_raw = goal_paths[ibest].q;
        res.goal_index = goal_paths[ibest].goal_index;
        res.start_index = goal_paths[ibest].start_index;
        res.waypoints_raw = (int)path_raw.size() / n;
        res.length_raw = goal_paths[ibest].length;
        if (prm.smoother_iterations > 0 && prm.post_processor == CC_POST_SHORTCUT_LINEAR) co_await shortcut_linear(path_raw, path);
        else if (prm.smoother_iterations > 0) co_await smooth(path_raw, path);
        else path = path_raw;
        res.waypoints = (int)path.size() / n;
        res.length = path_length(path);
        res.status = CC_PLAN_SOLVED;
        co_return 0;
    }
};

// ------------------------------------------------------------------------------------------------ the GPU checker
// All configurations of all groups in one cc_check_batch call (speculative: nothing stops early), then the first failing
// segment of every group in the reference's order.
int gpu_group_check(cc_planner* p, const double* cfg, int nconfigs, int ndof, const int32_t* group_begin, int ngroups,
                    const int32_t* segment_begin, int32_t* first_fail_segment, int32_t* fail_verdict, int32_t* first_fail_config)
{
    // stage through page-locked buffers: the copies of small batches then cost a few microseconds instead of tens
    const int cols = std::max(ndof, p->query_cols);
    const size_t need_q = (size_t)std::max(nconfigs, 1) * cols, need_v = (size_t)std::max(nconfigs, 1);
    if (need_q > p->pin_q_cap) {
        if (p->pin_q) cc_host_free(p->pin_q);
        p->pin_q = nullptr; p->pin_q_cap = 0;
        void* m = nullptr;
        const int rc = cc_host_alloc(need_q * 2 * sizeof(double), &m);
        if (rc != CC_OK) return rc;
        p->pin_q = (double*)m; p->pin_q_cap = need_q * 2;
    }
    if (need_v > p->pin_v_cap) {
        if (p->pin_v) cc_host_free(p->pin_v);
        p->pin_v = nullptr; p->pin_v_cap = 0;
        void* m = nullptr;
        const int rc = cc_host_alloc(need_v * 2 * sizeof(uint32_t), &m);
        if (rc != CC_OK) return rc;
        p->pin_v = (uint32_t*)m; p->pin_v_cap = need_v * 2;
    }
    uint32_t* verdict = p->pin_v;
    if (nconfigs > 0) {
        if (cols == ndof) std::memcpy(p->pin_q, cfg, (size_t)nconfigs * ndof * sizeof(double));
        else {
            // the body has mimic equations the device cannot hold: the library evaluates them per configuration on the host
            const int rce = cc_query_expand_rows(p->query, cfg, nconfigs, p->pin_q);
            if (rce != CC_OK) return rce;
        }
        const int rc = cc_check_batch(p->query, p->pin_q, nconfigs, CC_BATCH_ENV | CC_BATCH_SELF | CC_BATCH_FIRST_HIT, verdict, nullptr, nullptr, nullptr);
        if (rc != CC_OK) return rc;
    }
    for (int g = 0; g < ngroups; ++g) {
        first_fail_segment[g] = -1;
        fail_verdict[g] = 0;
        first_fail_config[g] = -1;
        for (int s = group_begin[g]; s < group_begin[g + 1] && first_fail_segment[g] < 0; ++s)
            for (int c = segment_begin[s]; c < segment_begin[s + 1]; ++c)
                if (verdict[c]) { first_fail_segment[g] = s - group_begin[g]; fail_verdict[g] = (int32_t)verdict[c]; first_fail_config[g] = c; break; }
    }
    return CC_OK;
}

}  // namespace

extern "C" {

int cc_planner_create(const cc_cspace* space, cc_planner** out)
{
    if (!space || !out || space->ndof <= 0 || !space->lower || !space->upper || !space->resolution)
        return fail(CC_ERR_INVALID, "cc_planner_create: bad configuration space");
    auto* p = new cc_planner();
    CSpace& cs = p->cs;
    cs.n = space->ndof;
    cs.lo.assign(space->lower, space->lower + cs.n);
    cs.hi.assign(space->upper, space->upper + cs.n);
    cs.res.assign(space->resolution, space->resolution + cs.n);
    cs.w2.assign(cs.n, 1.0);
    if (space->weights) for (int i = 0; i < cs.n; ++i) cs.w2[i] = space->weights[i] * space->weights[i];
    cs.circ.assign(cs.n, 0);
    if (space->circular) cs.circ.assign(space->circular, space->circular + cs.n);
    cs.any_circular = std::any_of(cs.circ.begin(), cs.circ.end(), [](uint8_t c) { return c != 0; });
    *out = p;
    return CC_OK;
}

int cc_planner_destroy(cc_planner* p)
{
    delete p;
    return CC_OK;
}

int cc_planner_set_query(cc_planner* p, cc_query* q)
{
    if (!p || !q) return fail(CC_ERR_INVALID, "cc_planner_set_query: NULL argument");
    cc_query_info info;
    const int rc = cc_query_get_info(q, &info);
    if (rc != CC_OK) return rc;
    if (info.ndof_active - info.nexternal != p->cs.n) return fail(CC_ERR_INVALID, "cc_planner_set_query: the query's active DOFs differ from the planner's");
    p->query = q;
    p->query_cols = info.ndof_active;
    p->fn = nullptr;
    return CC_OK;
}

int cc_planner_set_checker(cc_planner* p, cc_group_check_fn fn, void* user)
{
    if (!p || !fn) return fail(CC_ERR_INVALID, "cc_planner_set_checker: NULL argument");
    p->fn = fn;
    p->user = user;
    p->query = nullptr;
    return CC_OK;
}

int cc_planner_set_samplers(cc_planner* p, cc_config_sampler_fn goal_fn, cc_config_sampler_fn initial_fn, void* user)
{
    if (!p) return fail(CC_ERR_INVALID, "cc_planner_set_samplers: planner is NULL");
    p->sample_goal_fn = goal_fn;
    p->sample_initial_fn = initial_fn;
    p->sampler_user = user;
    return CC_OK;
}

int cc_planner_set_progress_callback(cc_planner* p, cc_progress_fn fn, void* user)
{
    if (!p) return fail(CC_ERR_INVALID, "cc_planner_set_progress_callback: planner is NULL");
    p->progress_fn = fn;
    p->progress_user = user;
    return CC_OK;
}

int cc_planner_edge_sample_times(cc_planner* p, const double* q0, const double* q1, int interval, double* out_times, int max_configs,
                                 int* out_count)
{
    if (!p || !q0 || !q1 || !out_count || interval < 0 || interval > 3) return fail(CC_ERR_INVALID, "cc_planner_edge_sample_times: bad arguments");
    std::vector<double> buf, times;
    const int c = edge_samples(p->cs, q0, q1, interval, buf, &times);
    *out_count = c;
    if (c > 0 && out_times) std::memcpy(out_times, times.data(), sizeof(double) * (size_t)std::min(c, std::max(max_configs, 0)));
    return CC_OK;
}

int cc_planner_edge_samples(cc_planner* p, const double* q0, const double* q1, int interval, double* out_configs, int max_configs,
                            int* out_count)
{
    if (!p || !q0 || !q1 || !out_count || interval < 0 || interval > 3) return fail(CC_ERR_INVALID, "cc_planner_edge_samples: bad arguments");
    std::vector<double> buf;
    const int c = edge_samples(p->cs, q0, q1, interval, buf);
    *out_count = c;
    if (c > 0 && out_configs) std::memcpy(out_configs, buf.data(), sizeof(double) * p->cs.n * (size_t)std::min(c, std::max(max_configs, 0)));
    return CC_OK;
}

int cc_planner_ramp_samples(cc_planner* p, const double* q0, const double* q1, const double* dq0, const double* dq1, double timeelapsed,
                            int interval, double* out_configs, double* out_times, uint8_t* out_filled, int max_configs, int* out_count)
{
    if (!p || !q0 || !q1 || !dq0 || !dq1 || !out_count || interval < 0 || interval > 3 || !(timeelapsed > 0))
        return fail(CC_ERR_INVALID, "cc_planner_ramp_samples: bad arguments");
    std::vector<double> buf, times;
    std::vector<uint8_t> filled;
    const int c = edge_samples_quadratic(p->cs, q0, q1, dq0, dq1, timeelapsed, interval, buf, &times, &filled);
    *out_count = c;
    const size_t m = c > 0 ? (size_t)std::min(c, std::max(max_configs, 0)) : 0;
    if (m && out_configs) std::memcpy(out_configs, buf.data(), sizeof(double) * p->cs.n * m);
    if (m && out_times) std::memcpy(out_times, times.data(), sizeof(double) * m);
    if (m && out_filled) std::memcpy(out_filled, filled.data(), m);
    return CC_OK;
}

// PlannerParameters::CheckPathAllConstraints for many edges at once (planner.h:373-379 -> DynamicsCollisionConstraint::Check,
// planningutils.cpp:2586-3629).  Every edge becomes one group; every sample one segment of 1 configuration, or of 3 with
// CFO_CheckWithPerturbation (the state, then +perturbation x resolution, then -perturbation x resolution, clamped to the
// limits, planningutils.cpp:2344-2366); the checker reports the first failing segment of each group, which is the sample the
// serial loop would have returned at.
int cc_planner_check_edges(cc_planner* p, int nedges, const double* q0, const double* q1, const double* dq0, const double* dq1,
                           const double* timeelapsed, int interval, int options, double perturbation, cc_edge_result* results,
                           double* invalid_values)
{
    if (!p || nedges < 0 || (nedges > 0 && (!q0 || !q1 || !results)) || interval < 0 || interval > 3)
        return fail(CC_ERR_INVALID, "cc_planner_check_edges: bad arguments");
    if (!p->query && !p->fn) return fail(CC_ERR_STATE, "cc_planner_check_edges: no checker bound (cc_planner_set_query)");
    if ((options & 3) != 3) return fail(CC_ERR_INVALID, "cc_planner_check_edges: CFO_CheckEnvCollisions | CFO_CheckSelfCollisions are checked together");
    const bool ramps = dq0 && dq1 && timeelapsed;
    const bool perturb = (options & 0x00010000) != 0 && perturbation > 0;      // CFO_CheckWithPerturbation (planner.h:37)
    const int n = p->cs.n, mult = perturb ? 3 : 1;
    const CSpace& cs = p->cs;
    std::vector<double> samples, times, q;
    std::vector<uint8_t> kinds;
    std::vector<int32_t> group_begin(1, 0), seg_begin(1, 0), sample_begin(1, 0);
    std::vector<double> all_samples, all_times;
    for (int e = 0; e < nedges; ++e) {
        samples.clear(); times.clear(); kinds.clear();
        const double* a = q0 + (size_t)e * n;
        const double* b = q1 + (size_t)e * n;
        int c;
        if (ramps && timeelapsed[e] > 0)
            c = edge_samples_quadratic(cs, a, b, dq0 + (size_t)e * n, dq1 + (size_t)e * n, timeelapsed[e], interval, samples, &times, &kinds);
        else c = edge_samples(cs, a, b, interval, samples, &times);
        cc_edge_result& r = results[e];
        std::memset(&r, 0, sizeof(r));
        r.first_invalid = -1;
        r.nsamples = c;
        if (c < 0) {                       // the walk itself failed: CFO_StateSettingError, nothing to check
            r.return_code = (int32_t)0x80000000u;
            r.nsamples = 0;
            c = 0;
            samples.clear(); times.clear();
        }
        for (int i = 0; i < c; ++i) {
            const double* sq = &samples[(size_t)i * n];
            q.insert(q.end(), sq, sq + n);
            if (perturb)
                for (double f : {perturbation, -perturbation})
                    for (int k = 0; k < n; ++k) q.push_back(std::min(std::max(sq[k] + f * cs.res[k], cs.lo[k]), cs.hi[k]));
            seg_begin.push_back((int32_t)(q.size() / n));
        }
        group_begin.push_back((int32_t)seg_begin.size() - 1);
        all_samples.insert(all_samples.end(), samples.begin(), samples.end());
        all_times.insert(all_times.end(), times.begin(), times.end());
        sample_begin.push_back((int32_t)all_times.size());
    }
    const int nconfigs = (int)(q.size() / n);
    std::vector<int32_t> ffs(std::max(nedges, 1), -1), fv(std::max(nedges, 1), 0), ffc(std::max(nedges, 1), -1);
    if (nconfigs > 0) {
        int rc;
        if (p->query) rc = gpu_group_check(p, q.data(), nconfigs, n, group_begin.data(), nedges, seg_begin.data(), ffs.data(), fv.data(), ffc.data());
        else rc = p->fn(p->user, q.data(), nconfigs, n, group_begin.data(), nedges, seg_begin.data(), (int)seg_begin.size() - 1, ffs.data(), fv.data(), ffc.data());
        if (rc != CC_OK) {
            if (!p->query) cc_internal_set_error("cc_planner_check_edges: the checker callback failed");
            return rc;
        }
    }
    for (int e = 0; e < nedges; ++e) {
        cc_edge_result& r = results[e];
        if (r.return_code != 0 || ffs[e] < 0) continue;
        const int i = ffs[e];
        r.first_invalid = i;
        r.return_code = (fv[e] & 1) ? 1 : 2;                                   // _CheckState: environment first, then self
        r.perturbed = perturb ? (ffc[e] - seg_begin[group_begin[e] + i]) : 0;  // 0 the state itself, 1 / 2 its +/- perturbation
        r.time_when_invalid = all_times[sample_begin[e] + i];
        if (invalid_values) std::memcpy(invalid_values + (size_t)e * n, &all_samples[((size_t)sample_begin[e] + i) * n], sizeof(double) * n);
    }
    (void)mult;
    return CC_OK;
}

}  // extern "C"

// kind 0: BiRRT; 1: BasicRRT; 2: BasicRRT with one-step extension (_nRRTExtentType == 1)
static int plan_batch(cc_planner* p, const cc_birrt_params* params, int kind, int nproblems, const double* starts, const double* goals,
                      const uint32_t* seeds, cc_plan_result* results, const int32_t* start_begin = nullptr, const int32_t* goal_begin = nullptr)
{
    if (!p || !params || nproblems < 0 || (nproblems > 0 && (!starts || !goals || !results)))
        return fail(CC_ERR_INVALID, "cc_birrt_plan_batch: bad arguments");
    if (!p->query && !p->fn) return fail(CC_ERR_STATE, "cc_birrt_plan_batch: no checker bound (cc_planner_set_query)");
    if (!(params->step_length > 0)) return fail(CC_ERR_INVALID, "cc_birrt_plan_batch: step_length must be > 0");
    const int n = p->cs.n;
    p->problems.clear();
    std::vector<Task<int>> tasks;
    tasks.reserve(nproblems);
    for (int k = 0; k < nproblems; ++k) {
        auto pr = std::make_unique<Problem>();
        pr->cs = &p->cs;
        pr->prm = *params;
        const int s0 = start_begin ? start_begin[k] : k, s1 = start_begin ? start_begin[k + 1] : k + 1;
        const int g0 = goal_begin ? goal_begin[k] : k, g1 = goal_begin ? goal_begin[k + 1] : k + 1;
        const bool goals_may_come = kind == 0 && p->sample_goal_fn != nullptr;       // BiRRT with a goal sampler may start without goals (rrt.h:349)
        if (s1 <= s0 || g1 < g0 || (g1 == g0 && !goals_may_come)) {
            p->problems.clear();
            return fail(CC_ERR_INVALID, "cc_birrt_plan_batch: a problem needs at least one initial and one goal configuration");
        }
        pr->starts_all.assign(starts + (size_t)s0 * n, starts + (size_t)s1 * n);
        pr->goals_all.assign(goals + (size_t)g0 * n, goals + (size_t)g1 * n);
        pr->start.assign(pr->starts_all.begin(), pr->starts_all.begin() + n);
        if (g1 > g0) pr->goal.assign(pr->goals_all.begin(), pr->goals_all.begin() + n);
        pr->seed = seeds ? seeds[k] : 0;
        pr->index = k;
        pr->sample_goal_fn = p->sample_goal_fn; pr->sample_initial_fn = p->sample_initial_fn; pr->sampler_user = p->sampler_user;
        pr->progress_fn = p->progress_fn; pr->progress_user = p->progress_user;
        pr->t_start = std::chrono::steady_clock::now();
        std::memset(&pr->res, 0, sizeof(pr->res));
        pr->res.start_index = pr->res.goal_index = -1;
        p->problems.push_back(std::move(pr));
    }
    for (int k = 0; k < nproblems; ++k) tasks.push_back(kind == 0 ? p->problems[k]->plan() : p->problems[k]->plan_basic(kind == 2));
    std::memset(p->totals, 0, sizeof(p->totals));

    // start every coroutine: each runs until its first request (or to completion)
    for (int k = 0; k < nproblems; ++k) {
        tasks[k].h.resume();
        if (tasks[k].h.done()) p->problems[k]->done = true;
    }
    std::vector<double> q;
    std::vector<int32_t> group_begin, seg_begin, ffs, fv, ffc;
    std::vector<int> owner;
    double t_check = 0, t_gather = 0, t_resume = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    for (;;) {
        const auto t0 = now();
        q.clear();
        group_begin.assign(1, 0);
        seg_begin.assign(1, 0);
        owner.clear();
        for (int k = 0; k < nproblems; ++k) {
            Problem& pr = *p->problems[k];
            if (pr.done || !pr.has_request) continue;
            const int base = (int)(q.size() / n);
            q.insert(q.end(), pr.group.q.begin(), pr.group.q.end());
            for (int s = 1; s <= pr.group.nseg(); ++s) seg_begin.push_back(base + pr.group.seg_begin[s]);
            group_begin.push_back((int32_t)seg_begin.size() - 1);
            owner.push_back(k);
        }
        if (owner.empty()) break;
        const int ngroups = (int)owner.size(), nconfigs = (int)(q.size() / n);
        ffs.assign(ngroups, -1);
        fv.assign(ngroups, 0);
        ffc.assign(ngroups, -1);
        const auto t1 = now();
        t_gather += secs(t0, t1);
        int rc;
        if (p->query) rc = gpu_group_check(p, q.data(), nconfigs, n, group_begin.data(), ngroups, seg_begin.data(), ffs.data(), fv.data(), ffc.data());
        else rc = p->fn(p->user, q.data(), nconfigs, n, group_begin.data(), ngroups, seg_begin.data(), (int)seg_begin.size() - 1, ffs.data(), fv.data(), ffc.data());
        if (rc != CC_OK) {
            if (!p->query) cc_internal_set_error("cc_birrt_plan_batch: the checker callback failed");
            p->problems.clear();
            return rc;
        }
        const auto t2 = now();
        t_check += secs(t1, t2);
        p->totals[0] += 1;
        p->totals[1] += 1;
        p->totals[2] += (uint64_t)nconfigs;
        for (int g = 0; g < ngroups; ++g) {
            Problem& pr = *p->problems[owner[g]];
            const int cbeg = seg_begin[group_begin[g]], cend = seg_begin[group_begin[g + 1]];
            int ff = ffs[g];
            uint64_t serial = (ff >= 0 && ffc[g] >= 0) ? (uint64_t)(ffc[g] - cbeg + 1) : (uint64_t)(cend - cbeg);
            // a segment that could not be sampled fails where it stands unless an earlier segment already did
            if (pr.group.state_error && (ff < 0 || ff >= pr.group.error_segment)) {
                ff = pr.group.error_segment;
                serial = (uint64_t)(seg_begin[group_begin[g] + ff] - cbeg);
            }
            pr.res.configs_submitted += (uint64_t)(cend - cbeg);
            pr.res.configs_serial += serial;
            pr.res.rounds++;
            p->totals[3] += serial;
            pr.reply.first_fail_segment = ff;
            pr.reply.verdict = fv[g];
            pr.has_request = false;
            auto h = pr.waiting;
            pr.waiting = nullptr;
            h.resume();
            if (tasks[owner[g]].h.done()) pr.done = true;
        }
        t_resume += secs(t2, now());
    }
    if (getenv("CC_DEBUG_PLANNER"))
        fprintf(stderr, "[planner] %llu rounds: gather %.2f ms, checker %.2f ms, planner host (resume) %.2f ms\n",
                (unsigned long long)p->totals[0], 1e3 * t_gather, 1e3 * t_check, 1e3 * t_resume);
    for (int k = 0; k < nproblems; ++k) results[k] = p->problems[k]->res;
    return CC_OK;
}

extern "C" {

int cc_birrt_plan_batch(cc_planner* p, const cc_birrt_params* params, int nproblems, const double* starts, const double* goals,
                        const uint32_t* seeds, cc_plan_result* results)
{
    return plan_batch(p, params, 0, nproblems, starts, goals, seeds, results);
}

int cc_birrt_plan_batch_multi(cc_planner* p, const cc_birrt_params* params, int nproblems, const int32_t* start_begin, const double* starts,
                              const int32_t* goal_begin, const double* goals, const uint32_t* seeds, cc_plan_result* results)
{
    if (nproblems > 0 && (!start_begin || !goal_begin)) return fail(CC_ERR_INVALID, "cc_birrt_plan_batch_multi: bad arguments");
    return plan_batch(p, params, 0, nproblems, starts, goals, seeds, results, start_begin, goal_begin);
}

int cc_basicrrt_plan_batch(cc_planner* p, const cc_birrt_params* params, int extent_type, int nproblems, const double* starts,
                           const double* goals, const uint32_t* seeds, cc_plan_result* results)
{
    return plan_batch(p, params, extent_type == 1 ? 2 : 1, nproblems, starts, goals, seeds, results);
}

int cc_planner_get_path(const cc_planner* p, int problem, int which, double* out, int max_waypoints, int* out_waypoints)
{
    if (!p || problem < 0 || problem >= (int)p->problems.size() || !out_waypoints) return fail(CC_ERR_INVALID, "cc_planner_get_path: bad arguments");
    const std::vector<double>& path = which == 1 ? p->problems[problem]->path_raw : p->problems[problem]->path;
    const int n = p->cs.n, w = (int)path.size() / n;
    *out_waypoints = w;
    if (out && max_waypoints > 0) std::memcpy(out, path.data(), sizeof(double) * n * (size_t)std::min(w, max_waypoints));
    return CC_OK;
}

int cc_planner_get_totals(const cc_planner* p, uint64_t* out4)
{
    if (!p || !out4) return fail(CC_ERR_INVALID, "cc_planner_get_totals: bad arguments");
    std::memcpy(out4, p->totals, sizeof(p->totals));
    return CC_OK;
}

}  // extern "C"
