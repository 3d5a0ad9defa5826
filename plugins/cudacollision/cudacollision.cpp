// cudacollision.cpp — OpenRAVE plugin `cudacollision`: a CollisionCheckerBase whose every verdict comes from
// libcudacollision.so (include/cudacollision.h), i.e. from hand-written sm_100a CUDA kernels on a B200.
//
//   env->SetCollisionChecker(RaveCreateCollisionChecker(env, "cudacollision"));          // normal plugin path
//   boost::dynamic_pointer_cast<CudaCollisionChecker>(env->GetCollisionChecker())->CheckCollisionBatch(...)
//   checker->SendCommand(sout, sinput)  with  "CheckBatch <robot> <n> <flags> <ptr q> <ptr verdict> <ptr firstpairs>"
//
// Built against the reference headers (include/openrave/*.h) inside an OpenRAVE source tree or against an installed
// OpenRAVE (CMakeLists.txt next to this file).  It cannot be LINKED or run in the B200 build container of this repository
// (libopenrave needs Boost >= 1.67, libxml2 and rapidjson), but it is type-checked there against the
// unmodified reference headers by tools/shim_syntax_check.sh (g++ -fsyntax-only with Boost stand-ins); the same object model
// is exercised through the Python twin openrave_b200/checker.py, which makes the identical sequence of C-ABI calls.
//
// Semantics kept (plugins/pqprave/collisionPQP.h): true = collision; report->Reset(options) first (:174-176); every
// colliding (link1,link2) appended in env-body / link / link order with minDistance = 0 (:476-485, :586-587); first
// hit only for self collision (:364-386); disabled links, attached bodies, excluded bodies/links skipped (:454-471,
// :555); CO_ActiveDOFs restricts links and self pairs (:373-377, :710-735); ray queries throw like pqprave's
// (:331-356); single-state queries read the bodies' CURRENT link transforms and never run FK (:438-447).
#include <openrave/openrave.h>
#include <openrave/plugin.h>
#include <openrave/plannerparameters.h>      // RRTParameters (cudabirrt)

#include <boost/bind/bind.hpp>
#include <boost/lexical_cast.hpp>
#include <boost/make_shared.hpp>

#include <cstring>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "cudacollision.h"

// the reference's plugins each carry this in their plugindefs.h (plugins/pqprave/plugindefs.h:30-47)
#ifndef FOREACHC
#define FOREACHC(it, v) for (auto it = (v).begin(); it != (v).end(); ++it)
#endif

using namespace OpenRAVE;

namespace {

// "a*x+b" family -> (a, b); returns false when the equation is not affine in exactly one joint variable.
// Mimic equations are fparser strings (kinbody.cpp:2308-2386); the WAM hand uses "JF1/3+0.8727", the PR2 gripper "0+x".
bool ParseAffineMimic(const std::string& eq, const std::string& var, double& a, double& b)
{
    // evaluate the expression symbolically as (coef, const) pairs with a tiny recursive-descent parser
    struct P {
        const std::string& s; const std::string& var; size_t i; bool ok;
        P(const std::string& s_, const std::string& v_) : s(s_), var(v_), i(0), ok(true) {}
        void ws() { while (i < s.size() && isspace((unsigned char)s[i])) ++i; }
        typedef std::pair<double, double> L;     // coef*x + const
        L atom() {
            ws();
            if (i < s.size() && s[i] == '(') { ++i; L v = expr(); ws(); if (i < s.size() && s[i] == ')') ++i; else ok = false; return v; }
            if (i < s.size() && (s[i] == '-' || s[i] == '+')) { bool neg = s[i] == '-'; ++i; L v = atom(); return neg ? L(-v.first, -v.second) : v; }
            if (s.compare(i, var.size(), var) == 0) { i += var.size(); return L(1.0, 0.0); }
            size_t used = 0;
            double v = 0;
            try { v = std::stod(s.substr(i), &used); } catch (...) { ok = false; return L(0, 0); }
            i += used;
            return L(0.0, v);
        }
        L term() {
            L v = atom();
            for (;;) {
                ws();
                if (i >= s.size() || (s[i] != '*' && s[i] != '/')) return v;
                char op = s[i++];
                L w = atom();
                if (op == '*') {
                    if (v.first != 0 && w.first != 0) { ok = false; return v; }
                    v = L(v.first * w.second + w.first * v.second, v.second * w.second);
                } else {
                    if (w.first != 0 || w.second == 0) { ok = false; return v; }
                    v = L(v.first / w.second, v.second / w.second);
                }
            }
        }
        L expr() {
            L v = term();
            for (;;) {
                ws();
                if (i >= s.size() || (s[i] != '+' && s[i] != '-')) return v;
                char op = s[i++];
                L w = term();
                v = op == '+' ? L(v.first + w.first, v.second + w.second) : L(v.first - w.first, v.second - w.second);
            }
        }
    } p(eq, var);
    P::L r = p.expr();
    p.ws();
    if (!p.ok || p.i != eq.size()) return false;
    a = r.first; b = r.second;
    return true;
}

inline void PackTransform(const Transform& t, double* o)
{
    o[0] = t.rot.x; o[1] = t.rot.y; o[2] = t.rot.z; o[3] = t.rot.w;       // RaveVector: x holds the scalar part (geometry.h:524)
    o[4] = t.trans.x; o[5] = t.trans.y; o[6] = t.trans.z; o[7] = 0;
}

}  // namespace

class CudaCollisionChecker : public CollisionCheckerBase
{
    struct BodyInfo : public UserData {
        KinBodyWeakPtr body;
        int id;                          // cc body id
        int transformStamp;              // KinBody::GetUpdateStamp at the last pose upload
        bool kinematicsDirty, pairsDirty, grabbedDirty;
        UserDataPtr cbGeometry, cbEnable, cbAttached, cbJoints, cbActive, cbGrabbed;
        std::map<int, std::vector<int32_t> > grabLists;   // grabbed body's cc id -> its (body, link) non-colliding list, computed when it was grabbed
        BodyInfo() : id(-1), transformStamp(-1), kinematicsDirty(true), pairsDirty(true), grabbedDirty(true) {}
    };
    typedef boost::shared_ptr<BodyInfo> BodyInfoPtr;

    struct QueryKey {
        int body; std::vector<int> active; int options;
        int affine;                      // RobotBase::GetAffineDOF() of a batched query over the robot's active DOFs, else 0
        double axis[3];
        QueryKey() : body(-1), options(0), affine(0) { axis[0] = axis[1] = 0; axis[2] = 1; }
        bool operator<(const QueryKey& r) const {
            if (body != r.body) return body < r.body;
            if (options != r.options) return options < r.options;
            if (affine != r.affine) return affine < r.affine;
            if (affine) for (int k = 0; k < 3; ++k) if (axis[k] != r.axis[k]) return axis[k] < r.axis[k];
            return active < r.active;
        }
    };

public:
    CudaCollisionChecker(EnvironmentBasePtr penv, std::istream& sinput) : CollisionCheckerBase(penv), _scene(NULL), _options(0), _device(0)
    {
        __description = ":Interface Author: cudacollision\n\nBatched collision checker for NVIDIA B200 (sm_100a): forward kinematics, "
                        "broadphase, BVH traversal and triangle tests run on the GPU for N configurations per call.";
        // creation arguments: "cudacollision [device | d0,d1,...]" — a device list replicates the scene on every listed GPU and
        // range-shards the batched entries over them inside the library (cc_scene_create_multi)
        std::string sdevices;
        sinput >> sdevices;
        std::vector<int> vdevices;
        {
            std::stringstream ss(sdevices);
            std::string tok;
            while (std::getline(ss, tok, ',')) if (!tok.empty()) vdevices.push_back(boost::lexical_cast<int>(tok));
        }
        if (vdevices.empty()) vdevices.push_back(0);
        _device = vdevices[0];
        _userdatakey = std::string("cudacollision") + boost::lexical_cast<std::string>(this);
        if (vdevices.size() > 1) _Check(cc_scene_create_multi(vdevices.data(), (int)vdevices.size(), &_scene));
        else _Check(cc_scene_create(_device, &_scene));
        RegisterCommand("CheckBatch", boost::bind(&CudaCollisionChecker::_CheckBatchCommand, this, boost::placeholders::_1, boost::placeholders::_2),
                        "CheckBatch robotname n flags qptr verdictptr firstpairsptr : batched FK + env + self check of n active-DOF "
                        "vectors (row-major doubles at qptr); writes uint32 verdicts (bit0 env, bit1 self) and int32[n][4] "
                        "{env link, env other body index, env other link, self pair packed i|j<<16} (-1 = none)");
        RegisterCommand("SetTriangleEpsilon", boost::bind(&CudaCollisionChecker::_SetTriEpsCommand, this, boost::placeholders::_1, boost::placeholders::_2),
                        "conservative margin (m) of the FP32 triangle test, default 0 = the reference's strict test");
    }

    virtual ~CudaCollisionChecker()
    {
        DestroyEnvironment();
        if (_scene) cc_scene_destroy(_scene);
    }

    // ------------------------------------------------------------------------------------------ options
    virtual bool SetCollisionOptions(int options)
    {
        if (options & CO_RayAnyHit) return false;   // rays are not provided (cf. fclcollision.cpp:171-180)
        if (options != _options) { _options = options; _DropQueries(); }
        return true;
    }
    virtual int GetCollisionOptions() const { return _options; }
    virtual void SetTolerance(dReal tolerance) { _tolerance = tolerance; _options |= CO_UseTolerance; }   // collisionPQP.h:426-428

    // ------------------------------------------------------------------------------------------ bodies
    virtual bool InitEnvironment()
    {
        std::vector<KinBodyPtr> vbodies;
        GetEnv()->GetBodies(vbodies);
        FOREACHC(it, vbodies) { _InitKinBody(*it); }
        return true;
    }

    virtual void DestroyEnvironment()
    {
        _DropQueries();
        std::vector<KinBodyPtr> vbodies;
        GetEnv()->GetBodies(vbodies);
        FOREACHC(it, vbodies) { RemoveKinBody(*it); }
    }

    virtual bool InitKinBody(KinBodyPtr pbody) { return !!_InitKinBody(pbody); }

    virtual void RemoveKinBody(KinBodyPtr pbody)
    {
        if (!pbody) return;
        BodyInfoPtr pinfo = boost::dynamic_pointer_cast<BodyInfo>(pbody->GetUserData(_userdatakey));
        if (!!pinfo) {
            cc_scene_remove_body(_scene, pinfo->id);
            _byid.erase(pinfo->id);
            pbody->RemoveUserData(_userdatakey);
            _DropQueries();
        }
    }

    /// environment clones create a fresh instance through the plugin and then call Clone (environment-core.h:3657-3662,
    /// pattern fclcollision.cpp:156-169): copy the options, the scene is rebuilt lazily from the cloned environment
    virtual void Clone(InterfaceBaseConstPtr preference, int cloningoptions)
    {
        CollisionCheckerBase::Clone(preference, cloningoptions);
        boost::shared_ptr<CudaCollisionChecker const> r = boost::dynamic_pointer_cast<CudaCollisionChecker const>(preference);
        if (!!r) { _options = r->_options; _DropQueries(); }
    }

    // ------------------------------------------------------------------------------------------ the batched entry (new)
    /// q: [n][ndof] active-DOF values of `probot` (RobotBase::SetActiveDOFValues semantics, robot.cpp:1193-1225); the other
    /// DOFs keep their current values.  flags: CC_BATCH_ENV | CC_BATCH_SELF.  firstpairs may be NULL.
    /// Replaces the per-sample SetDOFValues + CheckCollision + CheckSelfCollision loop of
    /// DynamicsCollisionConstraint::Check (planningutils.cpp:3405-3526).
    int CheckCollisionBatch(RobotBaseConstPtr probot, const dReal* q, int n, int ndof, uint32_t* verdict, int32_t* firstpairs, int flags)
    {
        if (ndof != probot->GetActiveDOF()) throw OPENRAVE_EXCEPTION_FORMAT0("cudacollision: ndof != active DOF count (joint values, then the affine values)", ORE_InvalidArguments);
        cc_query* query = _Bind(probot, &probot->GetActiveDOFIndices(), probot);
        if (_stale.erase(query)) _Check(cc_query_refresh(query));      // base pose / inactive DOF values of the robot moved
        std::vector<int32_t> fe(firstpairs ? n : 0), fs(firstpairs ? n : 0);
        cc_query_info info;
        cc_query_get_info(query, &info);
        if (info.ndof_active != ndof) {
            // the robot has mimic equations the device cannot hold (CC_MIMIC_EXTERNAL, _SyncKinematics): the library evaluates
            // them on the host and the rows grow by one column per such joint
            _vexpanded.resize((size_t)n * info.ndof_active);
            _Check(cc_query_expand_rows(query, q, n, _vexpanded.data()));
            q = _vexpanded.data();
        }
        _Check(cc_check_batch(query, q, n, flags, verdict, firstpairs ? fe.data() : NULL, firstpairs ? fs.data() : NULL, NULL));
        if (firstpairs) {
            std::vector<int32_t> quads(4 * (size_t)std::max(1, info.npairs_env + info.npairs_self));
            cc_query_get_pairs(query, quads.data());
            for (int i = 0; i < n; ++i) {
                int32_t* o = firstpairs + 4 * (size_t)i;
                o[0] = o[1] = o[2] = o[3] = -1;
                if (fe[i] >= 0) { o[0] = quads[4 * fe[i]]; o[1] = _EnvIndexOf(quads[4 * fe[i] + 1]); o[2] = quads[4 * fe[i] + 2]; }
                if (fs[i] >= 0) o[3] = quads[4 * fs[i]] | (quads[4 * fs[i] + 2] << 16);
            }
        }
        return 0;
    }

    /// the bound query behind CheckCollisionBatch for (robot, its active DOFs, current options): what cc_planner_set_query takes
    cc_query* BindQuery(RobotBaseConstPtr probot)
    {
        cc_query* query = _Bind(probot, &probot->GetActiveDOFIndices(), probot);
        if (_stale.erase(query)) _Check(cc_query_refresh(query));
        return query;
    }

    // ------------------------------------------------------------------------------------------ single-state overloads
    virtual bool CheckCollision(KinBodyConstPtr pbody, CollisionReportPtr report = CollisionReportPtr())
    {
        std::vector<KinBodyConstPtr> vbodyexcluded;
        std::vector<KinBody::LinkConstPtr> vlinkexcluded;
        return CheckCollision(pbody, vbodyexcluded, vlinkexcluded, report);
    }

    virtual bool CheckCollision(KinBodyConstPtr pbody, const std::vector<KinBodyConstPtr>& vbodyexcluded,
                                const std::vector<KinBody::LinkConstPtr>& vlinkexcluded, CollisionReportPtr report = CollisionReportPtr())
    {
        if (!!report) report->Reset(_options);
        // pqprave checks every body attached to pbody (collisionPQP.h:320-327) against every non-attached body
        std::set<KinBodyConstPtr> attached;
        pbody->GetAttached(attached);
        bool bcollision = false;
        FOREACHC(itatt, attached) {
            const bool bcommon = vbodyexcluded.empty() && vlinkexcluded.empty();
            if (bcommon && *itatt == pbody) {
                bcollision |= _CheckBoundEnv(pbody, report);      // the common case runs through the cached bound query
            } else if (bcommon && !!pbody->IsGrabbing(**itatt)) {
                continue;                                         // its links are query links of pbody's bound query (cc_body_set_grabbed)
            } else {
                bcollision |= _CheckPairs(*itatt, _EnvTriples(*itatt, attached, vbodyexcluded, vlinkexcluded, NULL), report);
            }
            if (bcollision && !report) return true;
        }
        return bcollision;
    }

    virtual bool CheckCollision(KinBodyConstPtr pbody1, KinBodyConstPtr pbody2, CollisionReportPtr report = CollisionReportPtr())
    {
        if (!!report) report->Reset(_options);
        if (pbody1->IsAttached(*pbody2)) return false;
        _Synchronize();
        std::vector<int32_t> triples;
        const int id2 = _Info(pbody2)->id;
        FOREACHC(it1, pbody1->GetLinks()) FOREACHC(it2, pbody2->GetLinks()) {
            triples.push_back((*it1)->GetIndex()); triples.push_back(id2); triples.push_back((*it2)->GetIndex());
        }
        return _CheckPairs(pbody1, triples, report);
    }

    virtual bool CheckCollision(KinBody::LinkConstPtr plink, CollisionReportPtr report = CollisionReportPtr())
    {
        std::vector<KinBodyConstPtr> vbodyexcluded;
        std::vector<KinBody::LinkConstPtr> vlinkexcluded;
        return CheckCollision(plink, vbodyexcluded, vlinkexcluded, report);
    }

    virtual bool CheckCollision(KinBody::LinkConstPtr plink1, KinBody::LinkConstPtr plink2, CollisionReportPtr report = CollisionReportPtr())
    {
        if (!!report) report->Reset(_options);
        _Synchronize();
        std::vector<int32_t> triples;
        triples.push_back(plink1->GetIndex()); triples.push_back(_Info(plink2->GetParent())->id); triples.push_back(plink2->GetIndex());
        return _CheckPairs(plink1->GetParent(), triples, report);
    }

    virtual bool CheckCollision(KinBody::LinkConstPtr plink, KinBodyConstPtr pbody, CollisionReportPtr report = CollisionReportPtr())
    {
        if (!!report) report->Reset(_options);
        if (plink->GetParent()->IsAttached(*pbody)) return false;
        _Synchronize();
        std::vector<int32_t> triples;
        const int id2 = _Info(pbody)->id;
        FOREACHC(it2, pbody->GetLinks()) { triples.push_back(plink->GetIndex()); triples.push_back(id2); triples.push_back((*it2)->GetIndex()); }
        return _CheckPairs(plink->GetParent(), triples, report);
    }

    virtual bool CheckCollision(KinBody::LinkConstPtr plink, const std::vector<KinBodyConstPtr>& vbodyexcluded,
                                const std::vector<KinBody::LinkConstPtr>& vlinkexcluded, CollisionReportPtr report = CollisionReportPtr())
    {
        if (!!report) report->Reset(_options);
        std::set<KinBodyConstPtr> attached;           // attached bodies are ignored for link queries (collisionchecker.h:250)
        attached.insert(plink->GetParent());
        const int only = plink->GetIndex();
        return _CheckPairs(plink->GetParent(), _EnvTriples(plink->GetParent(), attached, vbodyexcluded, vlinkexcluded, &only), report);
    }

    virtual bool CheckCollision(const RAY&, KinBody::LinkConstPtr, CollisionReportPtr = CollisionReportPtr())
    {
        throw OPENRAVE_EXCEPTION_FORMAT0("cudacollision: ray collision is not supported", ORE_NotImplemented);
    }
    virtual bool CheckCollision(const RAY&, KinBodyConstPtr, CollisionReportPtr = CollisionReportPtr())
    {
        throw OPENRAVE_EXCEPTION_FORMAT0("cudacollision: ray collision is not supported", ORE_NotImplemented);
    }
    virtual bool CheckCollision(const RAY&, CollisionReportPtr = CollisionReportPtr())
    {
        throw OPENRAVE_EXCEPTION_FORMAT0("cudacollision: ray collision is not supported", ORE_NotImplemented);
    }

    virtual bool CheckStandaloneSelfCollision(KinBodyConstPtr pbody, CollisionReportPtr report = CollisionReportPtr())
    {
        if (!!report) report->Reset(_options);
        if (pbody->GetLinks().size() <= 1) return false;
        const std::vector<int>* active = NULL;
        if (pbody->IsRobot() && (_options & CO_ActiveDOFs)) active = &RaveInterfaceConstCast<RobotBase>(pbody)->GetActiveDOFIndices();
        cc_query* query = _Bind(pbody, active);
        std::vector<double> T8 = _QueryPoses(query);
        uint32_t verdict = 0; int32_t fs = -1;
        _Check(cc_check_poses(query, T8.data(), 1, CC_BATCH_SELF, &verdict, NULL, &fs, NULL));
        // the body's own non-adjacent pairs come first in the pair order: a first hit that names a grabbed link belongs to
        // KinBody::CheckSelfCollision's grabbed loops (kinbodycollision.cpp:83-205), not to the standalone check
        bool bhit = false;
        if ((verdict & CC_VERDICT_SELF) && fs >= 0) {
            std::vector<int32_t> quads = _Quads(query);
            const int nl = (int)pbody->GetLinks().size();
            bhit = quads[4 * fs] < nl && quads[4 * fs + 2] < nl;
            if (bhit && !!report) {
                report->AddLinkCollision(*pbody->GetLinks().at(quads[4 * fs]), *pbody->GetLinks().at(quads[4 * fs + 2]));
                report->minDistance = 0;
            }
        }
        std::vector<KinBodyPtr> vgrabbed;
        pbody->GetGrabbed(vgrabbed);
        if (!bhit && !!report && (_options & CO_Distance) && vgrabbed.empty()) _FillDistance(query, T8, CC_BATCH_SELF, report);
        return bhit;
    }

    virtual bool CheckStandaloneSelfCollision(KinBody::LinkConstPtr plink, CollisionReportPtr report = CollisionReportPtr())
    {
        if (!!report) report->Reset(_options);
        KinBodyPtr pbody = plink->GetParent();
        _Synchronize();
        const int id = _Info(pbody)->id;
        std::vector<int32_t> triples;
        FOREACHC(itset, pbody->GetNonAdjacentLinks(KinBody::AO_Enabled)) {     // collisionPQP.h:388-417
            const int l0 = *itset & 0xffff, l1 = *itset >> 16;
            if (l0 == plink->GetIndex() || l1 == plink->GetIndex()) { triples.push_back(l0); triples.push_back(id); triples.push_back(l1); }
        }
        return _CheckPairs(pbody, triples, report, true);
    }

private:
    // ------------------------------------------------------------------------------------------ helpers
    void _Check(int rc) const
    {
        if (rc != CC_OK) throw OPENRAVE_EXCEPTION_FORMAT("cudacollision error %d: %s", rc % cc_last_error(), ORE_Failed);
    }

    BodyInfoPtr _Info(KinBodyConstPtr pbody) const
    {
        BodyInfoPtr pinfo = boost::dynamic_pointer_cast<BodyInfo>(pbody->GetUserData(_userdatakey));
        if (!pinfo) throw OPENRAVE_EXCEPTION_FORMAT("body %s is not known to cudacollision", pbody->GetName(), ORE_InvalidState);
        return pinfo;
    }

    int _EnvIndexOf(int ccbody) const
    {
        std::map<int, KinBodyWeakPtr>::const_iterator it = _byid.find(ccbody);
        KinBodyPtr p = it == _byid.end() ? KinBodyPtr() : it->second.lock();
        return !p ? -1 : p->GetEnvironmentBodyIndex();
    }

    BodyInfoPtr _InitKinBody(KinBodyConstPtr pbody)
    {
        BodyInfoPtr pinfo = boost::dynamic_pointer_cast<BodyInfo>(pbody->GetUserData(_userdatakey));
        if (!!pinfo && pinfo->body.lock() == pbody) return pinfo;     // clones can carry a stale pointer (collisionPQP.h:91-95)
        pinfo.reset(new BodyInfo());
        pinfo->body = boost::const_pointer_cast<KinBody>(pbody);
        _Check(cc_scene_add_body(_scene, pbody->GetName().c_str(), (int)pbody->GetLinks().size(), pbody->IsRobot() ? 1 : 0, &pinfo->id));
        _UploadMeshes(pbody, pinfo);
        boost::weak_ptr<BodyInfo> winfo(pinfo);
        pinfo->cbGeometry = pbody->RegisterChangeCallback(KinBody::Prop_LinkGeometry, boost::bind(&CudaCollisionChecker::_OnGeometry, this, winfo));
        pinfo->cbEnable = pbody->RegisterChangeCallback(KinBody::Prop_LinkEnable, boost::bind(&CudaCollisionChecker::_OnStructure, this, winfo, false));
        pinfo->cbAttached = pbody->RegisterChangeCallback(KinBody::Prop_BodyAttached, boost::bind(&CudaCollisionChecker::_OnStructure, this, winfo, false));
        pinfo->cbJoints = pbody->RegisterChangeCallback(KinBody::Prop_Joints | KinBody::Prop_LinkStatic, boost::bind(&CudaCollisionChecker::_OnStructure, this, winfo, true));
        pinfo->cbActive = pbody->RegisterChangeCallback(KinBody::Prop_RobotActiveDOFs, boost::bind(&CudaCollisionChecker::_OnStructure, this, winfo, false));
        pinfo->cbGrabbed = pbody->RegisterChangeCallback(KinBody::Prop_RobotGrabbed, boost::bind(&CudaCollisionChecker::_OnGrabbed, this, winfo));
        pbody->SetUserData(_userdatakey, pinfo);
        _byid[pinfo->id] = pinfo->body;
        _DropQueries();
        return pinfo;
    }

    void _UploadMeshes(KinBodyConstPtr pbody, BodyInfoPtr pinfo)
    {
        std::vector<double> verts;
        FOREACHC(itlink, pbody->GetLinks()) {
            const TriMesh& mesh = (*itlink)->GetCollisionData();                    // link-frame soup (kinbodylink.cpp:1202-1232)
            verts.resize(3 * mesh.vertices.size());
            for (size_t i = 0; i < mesh.vertices.size(); ++i) { verts[3 * i] = mesh.vertices[i].x; verts[3 * i + 1] = mesh.vertices[i].y; verts[3 * i + 2] = mesh.vertices[i].z; }
            _Check(cc_body_set_link_mesh(_scene, pinfo->id, (*itlink)->GetIndex(), (*itlink)->GetName().c_str(), verts.data(),
                                         (int)mesh.vertices.size(), mesh.indices.data(), (int)mesh.indices.size() / 3));
        }
        pinfo->transformStamp = -1;
    }

    void _OnGeometry(boost::weak_ptr<BodyInfo> winfo)
    {
        BodyInfoPtr pinfo = winfo.lock();
        KinBodyPtr pbody = !pinfo ? KinBodyPtr() : pinfo->body.lock();
        if (!!pbody) { _UploadMeshes(pbody, pinfo); pinfo->pairsDirty = true; _DropQueries(); }
    }

    /// Prop_RobotGrabbed fires inside Grab / Release: the non-colliding list of a NEWLY grabbed body is computed right here, at
    /// the poses of the grab — the state Grabbed::_listNonCollidingLinksWhenGrabbed refers to (the reference saves that state and
    /// evaluates lazily, kinbodygrab.cpp:63-200).  Lists of bodies grabbed earlier are kept, those of released bodies dropped.
    void _OnGrabbed(boost::weak_ptr<BodyInfo> winfo)
    {
        BodyInfoPtr pinfo = winfo.lock();
        if (!pinfo) return;
        pinfo->grabbedDirty = true;
        _DropQueries();
        KinBodyPtr pbody = pinfo->body.lock();
        if (!pbody) return;
        std::vector<KinBodyPtr> vgrabbed;
        pbody->GetGrabbed(vgrabbed);
        std::map<int, std::vector<int32_t> > kept;
        for (size_t g = 0; g < vgrabbed.size(); ++g) {
            BodyInfoPtr ginfo = _InitKinBody(vgrabbed[g]);
            std::map<int, std::vector<int32_t> >::iterator it = pinfo->grabLists.find(ginfo->id);
            if (it != pinfo->grabLists.end()) kept[ginfo->id] = it->second;
        }
        pinfo->grabLists.swap(kept);
        try {
            _Synchronize();
            for (size_t g = 0; g < vgrabbed.size(); ++g) {
                const int gid = _Info(vgrabbed[g])->id;
                if (pinfo->grabLists.find(gid) == pinfo->grabLists.end()) pinfo->grabLists[gid] = _NonCollidingList(pbody, pinfo, vgrabbed, g);
            }
        } catch (const std::exception& ex) {
            RAVELOG_WARN_FORMAT("cudacollision: non-colliding list of a grabbed body deferred to the next query (%s)", ex.what());
        }
    }

    /// Grabbed::_listNonCollidingLinksWhenGrabbed of vgrabbed[g] at the CURRENT poses: links of the grabber that are not rigidly
    /// attached to the grabbing link, not in the ignore set and not touching the grabbed body, plus the non-touching links of the
    /// bodies grabbed before it (kinbodygrab.cpp:63-200), as (cc body id, link index) pairs
    std::vector<int32_t> _NonCollidingList(KinBodyConstPtr pbody, BodyInfoPtr pinfo, const std::vector<KinBodyPtr>& vgrabbed, size_t g)
    {
        std::vector<int32_t> noncolliding;
        KinBodyPtr pgrabbed = vgrabbed[g];
        KinBody::LinkPtr plink = pbody->IsGrabbing(*pgrabbed);
        if (!plink) return noncolliding;
        BodyInfoPtr ginfo = _InitKinBody(pgrabbed);
        std::vector<KinBody::LinkPtr> vrigid;
        plink->GetRigidlyAttachedLinks(vrigid);
        std::list<KinBody::LinkConstPtr> listignored;
        pbody->GetIgnoredLinksOfGrabbed(pgrabbed, listignored);
        std::vector<std::pair<KinBodyConstPtr, BodyInfoPtr> > vsources(1, std::make_pair(pbody, pinfo));
        for (size_t o = 0; o < g; ++o) {
            KinBody::LinkPtr potherlink = pbody->IsGrabbing(*vgrabbed[o]);
            if (std::find(vrigid.begin(), vrigid.end(), potherlink) == vrigid.end()) vsources.push_back(std::make_pair(KinBodyConstPtr(vgrabbed[o]), _InitKinBody(vgrabbed[o])));
        }
        for (size_t isrc = 0; isrc < vsources.size(); ++isrc) {
            FOREACHC(itlink, vsources[isrc].first->GetLinks()) {
                if (isrc == 0 && (std::find(vrigid.begin(), vrigid.end(), *itlink) != vrigid.end() ||
                                  std::find(listignored.begin(), listignored.end(), KinBody::LinkConstPtr(*itlink)) != listignored.end())) continue;
                std::vector<int32_t> triples;
                for (size_t k = 0; k < pgrabbed->GetLinks().size(); ++k) { triples.push_back((*itlink)->GetIndex()); triples.push_back(ginfo->id); triples.push_back((int)k); }
                if (!_CheckPairs(vsources[isrc].first, triples, CollisionReportPtr(), true)) {
                    noncolliding.push_back(vsources[isrc].second->id);
                    noncolliding.push_back((*itlink)->GetIndex());
                }
            }
        }
        return noncolliding;
    }

    /// KinBody::_vGrabbedBodies -> cc_body_set_grabbed, with the lists _OnGrabbed computed at grab time (a list that could not
    /// be computed then — the checker was attached to the environment after the grab — is computed now)
    void _SyncGrabbed(KinBodyConstPtr pbody, BodyInfoPtr pinfo)
    {
        if (!pinfo->grabbedDirty) return;
        pinfo->grabbedDirty = false;
        _Check(cc_body_release_grabbed(_scene, pinfo->id, -1));
        std::vector<KinBodyPtr> vgrabbed;
        pbody->GetGrabbed(vgrabbed);
        for (size_t g = 0; g < vgrabbed.size(); ++g) {
            KinBodyPtr pgrabbed = vgrabbed[g];
            KinBody::LinkPtr plink = pbody->IsGrabbing(*pgrabbed);
            if (!plink) continue;
            BodyInfoPtr ginfo = _InitKinBody(pgrabbed);
            const Transform tinv = plink->GetTransform().inverse();
            std::vector<double> rel(8 * pgrabbed->GetLinks().size());
            for (size_t k = 0; k < pgrabbed->GetLinks().size(); ++k) PackTransform(tinv * pgrabbed->GetLinks()[k]->GetTransform(), &rel[8 * k]);
            if (pinfo->grabLists.find(ginfo->id) == pinfo->grabLists.end()) pinfo->grabLists[ginfo->id] = _NonCollidingList(pbody, pinfo, vgrabbed, g);
            const std::vector<int32_t>& noncolliding = pinfo->grabLists[ginfo->id];
            _Check(cc_body_set_grabbed(_scene, pinfo->id, ginfo->id, plink->GetIndex(), rel.data(), noncolliding.data(), (int)noncolliding.size() / 2));
            _SyncKinematics(pgrabbed, ginfo);      // grabbedBody.CheckSelfCollision needs its non-adjacent list
        }
    }

    /// link poses in the row layout of a bound query: the body's links, then those of its grabbed bodies (cc_query_get_links)
    std::vector<double> _QueryPoses(cc_query* query)
    {
        cc_query_info info;
        cc_query_get_info(query, &info);
        std::vector<int32_t> links(2 * (size_t)std::max(1, info.nlinks));
        cc_query_get_links(query, links.data());
        std::vector<double> T8(8 * (size_t)info.nlinks);
        for (int l = 0; l < info.nlinks; ++l) {
            KinBodyPtr p = _byid[links[2 * l]].lock();
            PackTransform(p->GetLinks().at(links[2 * l + 1])->GetTransform(), &T8[8 * (size_t)l]);
        }
        return T8;
    }

    KinBody::LinkConstPtr _QueryLink(cc_query* query, int qindex)
    {
        cc_query_info info;
        cc_query_get_info(query, &info);
        std::vector<int32_t> links(2 * (size_t)std::max(1, info.nlinks));
        cc_query_get_links(query, links.data());
        KinBodyPtr p = _byid[links[2 * qindex]].lock();
        return p->GetLinks().at(links[2 * qindex + 1]);
    }

    /// CO_Contacts: one CONTACT per intersecting triangle pair TriTriCollision accepts, depth 0 (collisionPQP.h:591-603)
    void _FillContacts(cc_query* query, const std::vector<double>& T8, int flags, CollisionReportPtr report)
    {
        const int maxcontacts = 1 << 16;
        std::vector<int32_t> pairs(maxcontacts), tris(2 * (size_t)maxcontacts);
        std::vector<uint8_t> has(maxcontacts);
        std::vector<double> posnorm(6 * (size_t)maxcontacts);
        int count = 0;
        _Check(cc_contacts_poses(query, T8.data(), flags, maxcontacts, &count, pairs.data(), tris.data(), has.data(), posnorm.data()));
        std::vector<int32_t> quads = _Quads(query);
        for (int i = 0; i < std::min(count, maxcontacts); ++i) {
            if (!has[i]) continue;
            const int p = pairs[i];
            KinBody::LinkConstPtr plink1 = _QueryLink(query, quads[4 * p]);
            KinBody::LinkConstPtr plink2 = quads[4 * p + 3] == 1 ? _QueryLink(query, quads[4 * p + 2])
                                                                    : KinBody::LinkConstPtr(_byid[quads[4 * p + 1]].lock()->GetLinks().at(quads[4 * p + 2]));
            const int icollision = report->AddLinkCollision(*plink1, *plink2);
            const double* o = &posnorm[6 * (size_t)i];
            report->vCollisionInfos[icollision].contacts.push_back(CONTACT(Vector(o[0], o[1], o[2]), Vector(o[3], o[4], o[5]), 0));
        }
    }

    /// CO_UseTolerance: numWithinTol = other bodies with at least one link pair within the tolerance (collisionPQP.h:443-497)
    void _FillTolerance(cc_query* query, const std::vector<double>& T8, const cc_query_info& info, CollisionReportPtr report)
    {
        std::vector<uint32_t> within(info.mask_words);
        _Check(cc_tolerance_poses(query, T8.data(), 1, CC_BATCH_ENV, _tolerance, within.data()));
        std::vector<int32_t> quads = _Quads(query);
        std::set<int> bodies;
        for (int p = 0; p < info.npairs_env; ++p) if ((within[p >> 5] >> (p & 31)) & 1u) bodies.insert(quads[4 * p + 1]);
        report->numWithinTol = (int)bodies.size();
    }

    /// CO_Distance on a bound single-state query (collisionPQP.h:622-648): minDistance, the closest pair, one contact
    void _FillDistance(cc_query* query, const std::vector<double>& T8, int flags, CollisionReportPtr report)
    {
        double dist = 0, pts[6];
        int32_t pair = -1;
        _Check(cc_distance_poses(query, T8.data(), 1, flags, 0.0, 0.0, 0.0, &dist, &pair, pts));
        if (pair < 0 || dist >= report->minDistance) return;
        std::vector<int32_t> quads = _Quads(query);
        KinBody::LinkConstPtr plink1 = _QueryLink(query, quads[4 * pair]);
        KinBody::LinkConstPtr plink2 = quads[4 * pair + 3] == 1 ? _QueryLink(query, quads[4 * pair + 2])
                                                                : KinBody::LinkConstPtr(_byid[quads[4 * pair + 1]].lock()->GetLinks().at(quads[4 * pair + 2]));
        report->minDistance = dist;
        const int icollision = report->AddLinkCollision(*plink1, *plink2);
        CollisionPairInfo& cpinfo = report->vCollisionInfos[icollision];
        cpinfo.contacts.resize(1);
        const Vector p1(pts[0], pts[1], pts[2]), p2(pts[3], pts[4], pts[5]);
        const dReal depth = RaveSqrt((p2 - p1).lengthsqr3());
        cpinfo.contacts[0].pos = p1;
        cpinfo.contacts[0].depth = -depth;
        cpinfo.contacts[0].norm = depth > 0 ? (p2 - p1) * (1 / depth) : Vector(0, 0, 0);
    }

    void _OnStructure(boost::weak_ptr<BodyInfo> winfo, bool kinematics)
    {
        BodyInfoPtr pinfo = winfo.lock();
        if (!pinfo) return;
        pinfo->transformStamp = -1;
        pinfo->pairsDirty = true;
        if (kinematics) pinfo->kinematicsDirty = true;
        _DropQueries();
    }

    void _DropQueries()
    {
        for (std::map<QueryKey, cc_query*>::iterator it = _queries.begin(); it != _queries.end(); ++it) cc_query_destroy(it->second);
        _queries.clear();
        _stale.clear();
    }

    /// push poses / enable flags / attachments of every body whose update stamp moved, then commit (idempotent)
    void _Synchronize()
    {
        std::vector<KinBodyPtr> vbodies;
        GetEnv()->GetBodies(vbodies);
        bool changed = false;
        std::set<int> moved;                 // cc ids of the bodies whose poses were pushed in this pass
        std::vector<double> T8;
        std::vector<uint8_t> enabled;
        FOREACHC(itbody, vbodies) {
            BodyInfoPtr pinfo = _InitKinBody(*itbody);
            if (pinfo->transformStamp == (*itbody)->GetUpdateStamp()) continue;
            moved.insert(pinfo->id);
            T8 = _Poses(*itbody);
            _Check(cc_body_set_link_transforms(_scene, pinfo->id, T8.data()));
            enabled.resize((*itbody)->GetLinks().size());
            for (size_t i = 0; i < enabled.size(); ++i) enabled[i] = (*itbody)->GetLinks()[i]->IsEnabled() ? 1 : 0;
            _Check(cc_body_set_link_enabled(_scene, pinfo->id, enabled.data()));
            if ((*itbody)->GetDOF() > 0 && !pinfo->kinematicsDirty) {
                std::vector<dReal> values;
                (*itbody)->GetDOFValues(values);
                _Check(cc_body_set_dof_values(_scene, pinfo->id, values.data(), (int)values.size()));
            }
            pinfo->transformStamp = (*itbody)->GetUpdateStamp();
            changed = true;
        }
        if (changed) {
            // attachments (KinBody::IsAttached: grabbed / attached bodies are never checked against each other)
            for (size_t i = 0; i < vbodies.size(); ++i)
                for (size_t j = i + 1; j < vbodies.size(); ++j)
                    _Check(cc_body_set_attached(_scene, _Info(vbodies[i])->id, _Info(vbodies[j])->id, vbodies[i]->IsAttached(*vbodies[j]) ? 1 : 0));
        }
        _Check(cc_scene_commit(_scene));
        // Bound queries captured poses (static slots, base, inactive DOF values, hoisted verdicts).  Pose changes do not change
        // a pair list, so the queries are refreshed, not re-created: at once when another body moved, lazily (before the next
        // FK-mode call) when only the query's own body or what it grabs moved — the single-state overloads pass that body's
        // transforms in anyway.  Structural changes (enable flags, attachments, grabs, joints, geometry, active DOFs) arrive
        // through the change callbacks and drop the queries.
        for (std::map<QueryKey, cc_query*>::iterator it = _queries.begin(); it != _queries.end() && !moved.empty(); ++it) {
            std::set<int> own;
            own.insert(it->first.body);
            KinBodyPtr pown = _byid[it->first.body].lock();
            if (!!pown) {
                std::vector<KinBodyPtr> vgrabbed;
                pown->GetGrabbed(vgrabbed);
                FOREACHC(itg, vgrabbed) { own.insert(_Info(*itg)->id); }
            }
            bool others = false;
            FOREACHC(itm, moved) { if (own.find(*itm) == own.end()) { others = true; break; } }
            if (others) { _Check(cc_query_refresh(it->second)); _stale.erase(it->second); }
            else _stale.insert(it->second);
        }
    }

    std::vector<double> _Poses(KinBodyConstPtr pbody) const
    {
        std::vector<Transform> vtrans;
        pbody->GetLinkTransformations(vtrans);
        std::vector<double> T8(8 * vtrans.size());
        for (size_t i = 0; i < vtrans.size(); ++i) PackTransform(vtrans[i], &T8[8 * i]);
        return T8;
    }

    /// FK chain program read off the computed KinBody (kinbody.cpp:2286-2481; accessors kinbodyjoint.cpp:1567-1582)
    void _SyncKinematics(KinBodyConstPtr pbody, BodyInfoPtr pinfo)
    {
        if (pinfo->kinematicsDirty) {
            const std::vector<KinBody::JointPtr>& vjoints = pbody->GetDependencyOrderedJointsAll();
            std::vector<cc_joint> joints(vjoints.size());
            std::vector<int> vexternal;
            for (size_t k = 0; k < vjoints.size(); ++k) {
                const KinBody::Joint& j = *vjoints[k];
                cc_joint& o = joints[k];
                memset(&o, 0, sizeof(o));
                if (j.GetDOF() != 1) throw OPENRAVE_EXCEPTION_FORMAT("cudacollision: joint %s has %d DOF (1-DOF joints only)", j.GetName() % j.GetDOF(), ORE_NotImplemented);
                o.type = j.IsRevolute(0) ? CC_JOINT_REVOLUTE : CC_JOINT_PRISMATIC;
                o.parent_link = !j.GetHierarchyParentLink() ? 0 : j.GetHierarchyParentLink()->GetIndex();
                o.child_link = j.GetHierarchyChildLink()->GetIndex();
                o.dof_index = j.GetDOFIndex();
                o.mimic_dof = -1;
                o.is_static = j.IsStatic() ? 1 : 0;
                o.is_circular = j.IsCircular(0) ? 1 : 0;
                std::vector<dReal> lo, hi;
                j.GetLimits(lo, hi);
                o.lower = lo.at(0); o.upper = hi.at(0);
                PackTransform(j.GetInternalHierarchyLeftTransform(), o.left);
                PackTransform(j.GetInternalHierarchyRightTransform(), o.right);
                const Vector& ax = j.GetInternalHierarchyAxis(0);
                o.axis[0] = ax.x; o.axis[1] = ax.y; o.axis[2] = ax.z;
                if (j.IsMimic(0) && !j.IsStatic()) {
                    std::vector<int> vmimicdofs;
                    j.GetMimicDOFIndices(vmimicdofs, 0);
                    o.dof_index = -1;
                    if (vmimicdofs.size() == 1 && ParseAffineMimic(j.GetMimicEquation(0, 0), pbody->GetJointFromDOFIndex(vmimicdofs[0])->GetName(), o.mimic_a, o.mimic_b))
                        o.mimic_dof = vmimicdofs[0];
                    else {
                        // any other fparser expression (kinbody.cpp:2308-2320): the library compiles the text and evaluates it on
                        // the host per configuration (cc_body_set_mimic_equation / cc_query_expand_rows)
                        o.mimic_dof = CC_MIMIC_EXTERNAL;
                        o.reserved = (int)vexternal.size();
                        vexternal.push_back((int)k);
                    }
                }
            }
            const int ndof = pbody->GetDOF(), nl = (int)pbody->GetLinks().size();
            std::vector<uint8_t> affects((size_t)std::max(1, ndof) * nl, 0);
            for (int d = 0; d < ndof; ++d) {
                const int jointindex = pbody->GetJointFromDOFIndex(d)->GetJointIndex();
                for (int l = 0; l < nl; ++l) affects[(size_t)d * nl + l] = pbody->DoesAffect(jointindex, l) ? 1 : 0;
            }
            _Check(cc_body_set_kinematics(_scene, pinfo->id, ndof, (int)joints.size(), joints.data(), affects.data()));
            if (!vexternal.empty()) {
                std::vector<const char*> vnames(vjoints.size());
                for (size_t k = 0; k < vjoints.size(); ++k) vnames[k] = vjoints[k]->GetName().c_str();
                for (size_t e = 0; e < vexternal.size(); ++e) {
                    const std::string eq = vjoints[vexternal[e]]->GetMimicEquation(0, 0);
                    if (cc_body_set_mimic_equation(_scene, pinfo->id, vexternal[e], eq.c_str(), vnames.data(), (int)vnames.size()) != CC_OK)
                        throw OPENRAVE_EXCEPTION_FORMAT("cudacollision: joint %s: %s", vjoints[vexternal[e]]->GetName() % cc_last_error(), ORE_NotImplemented);
                }
            }
            std::vector<dReal> values;
            pbody->GetDOFValues(values);
            if (ndof > 0) _Check(cc_body_set_dof_values(_scene, pinfo->id, values.data(), ndof));
            pinfo->kinematicsDirty = false;
        }
        if (pinfo->pairsDirty) {
            // KinBody::GetNonAdjacentLinks probes the initial pose through CheckCollision(link,link) of THIS checker
            // (kinbody.cpp:5563-5577), i.e. through the GPU; the list arrives in CompareNonAdjacentFarthest order.
            const std::vector<int>& pairs = pbody->GetNonAdjacentLinks(0);
            std::vector<int32_t> packed(pairs.begin(), pairs.end());
            _Check(cc_body_set_nonadjacent_pairs(_scene, pinfo->id, packed.data(), (int)packed.size()));
            pinfo->pairsDirty = false;
        }
    }

    /// paffinerobot: bind over the robot's affine active DOFs too (RobotBase::SetActiveDOFs(indices, affine, axis), robot.cpp:1193-1225)
    cc_query* _Bind(KinBodyConstPtr pbody, const std::vector<int>* active, RobotBaseConstPtr paffinerobot = RobotBaseConstPtr())
    {
        _Synchronize();
        BodyInfoPtr pinfo = _Info(pbody);
        _SyncKinematics(pbody, pinfo);
        _SyncGrabbed(pbody, pinfo);
        _Check(cc_scene_commit(_scene));
        QueryKey key;
        key.body = pinfo->id; key.options = _options;
        if (active) key.active = *active;
        if (!!paffinerobot && paffinerobot->GetAffineDOF() != 0) {
            key.affine = paffinerobot->GetAffineDOF();
            const Vector& ax = paffinerobot->GetAffineRotationAxis();
            key.axis[0] = ax.x; key.axis[1] = ax.y; key.axis[2] = ax.z;
        }
        std::map<QueryKey, cc_query*>::iterator it = _queries.find(key);
        if (it != _queries.end()) return it->second;
        cc_query* query = NULL;
        std::vector<int32_t> act(key.active.begin(), key.active.end());
        if (key.affine) {
            // _vRotationQuatLimitStart = GetAffineRotationQuatLimits() / its norm (robot.h:1097-1099, robot.cpp:1031)
            Vector qs = paffinerobot->GetAffineRotationQuatLimits();
            const dReal len = RaveSqrt(qs.lengthsqr4());
            double qstart[4] = {1, 0, 0, 0};
            if (len > 0) { qstart[0] = qs.x / len; qstart[1] = qs.y / len; qstart[2] = qs.z / len; qstart[3] = qs.w / len; }
            int32_t none = 0;
            _Check(cc_query_create_affine(_scene, pinfo->id, act.empty() ? &none : act.data(), (int)act.size(), key.affine, key.axis, qstart, _options, &query));
        } else
        _Check(cc_query_create(_scene, pinfo->id, active ? act.data() : NULL, (int)act.size(), _options, &query));
        _queries[key] = query;
        return query;
    }

    std::vector<int32_t> _Quads(cc_query* query) const
    {
        cc_query_info info;
        cc_query_get_info(query, &info);
        std::vector<int32_t> quads(4 * (size_t)std::max(1, info.npairs_env + info.npairs_self));
        cc_query_get_pairs(query, quads.data());
        return quads;
    }

    /// env check of one body through its cached bound query, current poses, complete pair mask when a report is wanted
    bool _CheckBoundEnv(KinBodyConstPtr pbody, CollisionReportPtr report)
    {
        const std::vector<int>* active = NULL;
        if (pbody->IsRobot() && (_options & CO_ActiveDOFs)) active = &RaveInterfaceConstCast<RobotBase>(pbody)->GetActiveDOFIndices();
        cc_query* query = _Bind(pbody, active);
        cc_query_info info;
        cc_query_get_info(query, &info);
        std::vector<double> T8 = _QueryPoses(query);       // grabbed links included: the bound query covers them
        const bool wantall = !!report || _UseCallbacks();       // a callback may veto pairs: then every colliding pair is needed
        std::vector<uint32_t> mask(!wantall ? 0 : info.mask_words);
        uint32_t verdict = 0;
        _Check(cc_check_poses(query, T8.data(), 1, CC_BATCH_ENV | (!wantall ? 0 : CC_BATCH_ALL_PAIRS), &verdict, NULL, NULL, !wantall ? NULL : mask.data()));
        if ((verdict & CC_VERDICT_ENV) && wantall && _FillReport(pbody, query, mask, info.npairs_env, report) == 0) verdict &= ~(uint32_t)CC_VERDICT_ENV;
        if ((verdict & CC_VERDICT_ENV) && !!report && (_options & CO_Contacts)) _FillContacts(query, T8, CC_BATCH_ENV, report);
        if (!(verdict & CC_VERDICT_ENV) && !!report && (_options & CO_Distance)) _FillDistance(query, T8, CC_BATCH_ENV, report);
        if (!!report && (_options & CO_UseTolerance)) _FillTolerance(query, T8, info, report);
        return (verdict & CC_VERDICT_ENV) != 0;
    }

    /// collision callbacks are consulted unless CO_IgnoreCallbacks is set (collisionPQP.h:569)
    bool _UseCallbacks() const { return !(_options & CO_IgnoreCallbacks) && GetEnv()->HasRegisteredCollisionCallbacks(); }

    /// a colliding link pair stands unless a registered callback answers anything but CA_DefaultAction for a report that holds
    /// just this pair (collisionPQP.h:606-617; fclcollision.cpp:1002-1011)
    bool _PairStands(const KinBody::Link& link1, const KinBody::Link& link2)
    {
        if (!_UseCallbacks()) return true;
        CollisionReportPtr preport(new CollisionReport());
        preport->Reset(_options);
        preport->AddLinkCollision(link1, link2);
        preport->minDistance = 0;
        std::list<EnvironmentBase::CollisionCallbackFn> listcallbacks;
        GetEnv()->GetRegisteredCollisionCallbacks(listcallbacks);
        FOREACHC(itfn, listcallbacks) {
            if ((*itfn)(preport, false) != CA_DefaultAction) return false;
        }
        return true;
    }

    /// every colliding pair of the mask that the callbacks let stand goes into the report (may be null); returns how many stood
    int _FillReport(KinBodyConstPtr pbody, cc_query* query, const std::vector<uint32_t>& mask, int npairs, CollisionReportPtr report)
    {
        std::vector<int32_t> quads = _Quads(query);
        int nstanding = 0;
        for (int p = 0; p < npairs; ++p) {
            if (!((mask[p >> 5] >> (p & 31)) & 1u)) continue;
            KinBodyPtr pother = _byid[quads[4 * p + 1]].lock();
            if (!pother) continue;
            KinBody::LinkConstPtr plink1 = _QueryLink(query, quads[4 * p]);
            KinBody::LinkConstPtr plink2 = pother->GetLinks().at(quads[4 * p + 2]);
            if (!_PairStands(*plink1, *plink2)) continue;
            ++nstanding;
            if (!!report) report->AddLinkCollision(*plink1, *plink2);                                                 // collisionPQP.h:586
        }
        if (!!report && nstanding > 0) report->minDistance = 0;                                                       // :587
        return nstanding;
    }

    /// (link, other body, other link) triples of pbody against the environment in pqprave's order (collisionPQP.h:448-487)
    std::vector<int32_t> _EnvTriples(KinBodyConstPtr pbody, const std::set<KinBodyConstPtr>& attached, const std::vector<KinBodyConstPtr>& vbodyexcluded,
                                     const std::vector<KinBody::LinkConstPtr>& vlinkexcluded, const int* onlylink)
    {
        _Synchronize();
        std::vector<KinBodyPtr> vbodies;
        GetEnv()->GetBodies(vbodies);
        std::vector<int32_t> triples;
        const bool activeonly = pbody->IsRobot() && (_options & CO_ActiveDOFs) && !onlylink;
        RobotBaseConstPtr probot = activeonly ? RaveInterfaceConstCast<RobotBase>(pbody) : RobotBaseConstPtr();
        FOREACHC(itbody, vbodies) {
            if (attached.find(*itbody) != attached.end()) continue;
            if (std::find(vbodyexcluded.begin(), vbodyexcluded.end(), *itbody) != vbodyexcluded.end()) continue;
            const int id2 = _Info(*itbody)->id;
            FOREACHC(itlink, pbody->GetLinks()) {
                if (onlylink && (*itlink)->GetIndex() != *onlylink) continue;
                if (std::find(vlinkexcluded.begin(), vlinkexcluded.end(), *itlink) != vlinkexcluded.end()) continue;
                if (activeonly) {                                           // _IsActiveLink (collisionPQP.h:710-735)
                    bool affected = false;
                    FOREACHC(itdof, probot->GetActiveDOFIndices())
                        if (probot->DoesAffect(probot->GetJointFromDOFIndex(*itdof)->GetJointIndex(), (*itlink)->GetIndex())) { affected = true; break; }
                    if (!affected) continue;
                }
                FOREACHC(itlink2, (*itbody)->GetLinks()) {
                    if (std::find(vlinkexcluded.begin(), vlinkexcluded.end(), *itlink2) != vlinkexcluded.end()) continue;
                    triples.push_back((*itlink)->GetIndex()); triples.push_back(id2); triples.push_back((*itlink2)->GetIndex());
                }
            }
        }
        return triples;
    }

    /// explicit pair list at the bodies' current poses (disabled links never collide: the C ABI keeps their index and skips them)
    bool _CheckPairs(KinBodyConstPtr pbody, const std::vector<int32_t>& triples, CollisionReportPtr report, bool firstonly = false)
    {
        if (triples.empty()) return false;
        BodyInfoPtr pinfo = _Info(pbody);
        cc_query* query = NULL;
        _Check(cc_query_create_pairs(_scene, pinfo->id, triples.data(), (int)triples.size() / 3, &query));
        cc_query_info info;
        cc_query_get_info(query, &info);
        std::vector<double> T8 = _Poses(pbody);
        std::vector<uint32_t> mask(info.mask_words);
        uint32_t verdict = 0; int32_t fe = -1, fs = -1;
        int rc = cc_check_poses(query, T8.data(), 1, CC_BATCH_ENV | CC_BATCH_SELF | (firstonly ? 0 : CC_BATCH_ALL_PAIRS), &verdict, &fe, &fs, firstonly ? NULL : mask.data());
        if (rc == CC_OK && verdict && firstonly) {
            std::vector<int32_t> quads = _Quads(query);
            const int p = fs >= 0 ? fs : fe;
            KinBodyPtr pother = _byid[quads[4 * p + 1]].lock();
            if (!!pother && !!report) report->AddLinkCollision(*pbody->GetLinks().at(quads[4 * p]), *pother->GetLinks().at(quads[4 * p + 2]));
            if (!!report) report->minDistance = 0;
        } else if (rc == CC_OK && verdict && (!!report || _UseCallbacks())) {
            if (_FillReport(pbody, query, mask, info.npairs_env + info.npairs_self, report) == 0) verdict = 0;        // every pair vetoed
        }
        cc_query_destroy(query);
        _Check(rc);
        return verdict != 0;
    }

    bool _CheckBatchCommand(std::ostream& sout, std::istream& sinput)
    {
        std::string robotname;
        int n = 0, flags = 0;
        uintptr_t pq = 0, pverdict = 0, ppairs = 0;
        sinput >> robotname >> n >> flags >> pq >> pverdict >> ppairs;
        RobotBasePtr probot = GetEnv()->GetRobot(robotname);
        if (!probot || !sinput) return false;
        CheckCollisionBatch(probot, (const dReal*)pq, n, probot->GetActiveDOF(), (uint32_t*)pverdict, (int32_t*)ppairs, flags);
        sout << n;
        return true;
    }

    bool _SetTriEpsCommand(std::ostream&, std::istream& sinput)
    {
        double eps = 0;
        sinput >> eps;
        return !!sinput && cc_scene_set_tri_epsilon(_scene, eps) == CC_OK;
    }

    cc_scene* _scene;
    int _options, _device;
    dReal _tolerance = 0;
    std::string _userdatakey;
    std::map<int, KinBodyWeakPtr> _byid;
    std::map<QueryKey, cc_query*> _queries;
    std::set<cc_query*> _stale;            // queries whose own body moved since their last refresh (matters to the FK-mode entry only)
    std::vector<double> _vexpanded;        // rows of CheckCollisionBatch with the host-evaluated mimic columns filled in
};

#include "cudabatched.h"

// ---------------------------------------------------------------------------------------------- planner: "cudabirrt"
// RaveCreatePlanner(env, "cudabirrt"): BirrtPlanner's interface (plugins/rplanners/rrt.h:279-659) on cc_birrt_plan_batch.  InitPlan reads
// the same RRTParameters (goal bias is BirrtPlanner's constant 0.01, rrt.h:315), PlanPath plans vinitialconfig -> vgoalconfig and
// appends the smoothed waypoints to the trajectory; PlanPathBatch is the new entry for callers that hold many queries.  Needs the
// environment's collision checker to be "cudacollision".  Several initial / goal configurations, _minimumgoalpaths and the goal /
// initial samplers (_samplegoalfn, _sampleinitialfn), the registered progress callbacks and _nMaxPlanningTime are carried over.
class CudaBirrtPlanner : public PlannerBase
{
public:
    CudaBirrtPlanner(EnvironmentBasePtr penv) : PlannerBase(penv), _planner(NULL)
    {
        __description = ":Interface Author: cudacollision\n\nBi-directional RRT (RRT-Connect) + linear shortcutting with every edge check batched on the GPU; "
                        "many planning problems advance in lockstep.";
    }
    virtual ~CudaBirrtPlanner() { if (_planner) cc_planner_destroy(_planner); }

    virtual PlannerStatus InitPlan(RobotBasePtr probot, PlannerParametersConstPtr pparams)
    {
        EnvironmentLock lock(GetEnv()->GetMutex());
        _parameters.reset(new RRTParameters());
        _parameters->copy(pparams);
        _robot = probot;
        const int ndof = _parameters->GetDOF();
        if ((int)probot->GetActiveDOFIndices().size() != ndof) return PlannerStatus("cudabirrt plans over the robot's active joint DOFs", PS_Failed);
        // PlannerParameters::SetRobotActiveJoints' configuration space (planner.cpp:650-697)
        std::vector<dReal> weights;
        probot->GetActiveDOFWeights(weights);
        std::vector<uint8_t> circular(ndof, 0);
        for (int i = 0; i < ndof; ++i) {
            KinBody::JointPtr pjoint = probot->GetJointFromDOFIndex(probot->GetActiveDOFIndices()[i]);
            circular[i] = pjoint->IsCircular(probot->GetActiveDOFIndices()[i] - pjoint->GetDOFIndex()) ? 1 : 0;
        }
        cc_cspace space = {ndof, 0, _parameters->_vConfigLowerLimit.data(), _parameters->_vConfigUpperLimit.data(),
                           _parameters->_vConfigResolution.data(), weights.data(), circular.data()};
        if (_planner) { cc_planner_destroy(_planner); _planner = NULL; }
        if (cc_planner_create(&space, &_planner) != CC_OK) return PlannerStatus(cc_last_error(), PS_Failed);
        boost::shared_ptr<CudaCollisionChecker> pchecker = boost::dynamic_pointer_cast<CudaCollisionChecker>(GetEnv()->GetCollisionChecker());
        if (!pchecker) return PlannerStatus("cudabirrt needs the cudacollision checker (env->SetCollisionChecker)", PS_Failed);
        // the planners run with CO_ActiveDOFs (rrt.h:61, 387)
        CollisionOptionsStateSaver optionstate(pchecker, pchecker->GetCollisionOptions() | CO_ActiveDOFs, false);
        if (cc_planner_set_query(_planner, pchecker->BindQuery(probot)) != CC_OK) return PlannerStatus(cc_last_error(), PS_Failed);
        memset(&_birrt, 0, sizeof(_birrt));
        _birrt.step_length = _parameters->_fStepLength > 0 ? _parameters->_fStepLength : 0.04;
        _birrt.goal_bias = 0.01;
        _birrt.point_tolerance = 0.2;
        _birrt.max_iterations = _parameters->_nMaxIterations > 0 ? _parameters->_nMaxIterations : 10000;
        _birrt.path_optimize_iterations = 10;
        _birrt.smoother_iterations = _parameters->_sPostProcessingPlanner.size() > 0 ? 100 : 0;     // linearsmoother's default budget
        _birrt.max_extend_steps = 100;
        _birrt.speculate_steps = -8;
        _birrt.post_processor = _parameters->_sPostProcessingPlanner == "shortcut_linear" ? CC_POST_SHORTCUT_LINEAR
                                : _parameters->_sPostProcessingPlanner == "linearsmoother 0" ? CC_POST_LINEARSMOOTHER_SHORTCUTS   // creation args (linearsmoother.cpp:25-26)
                                : CC_POST_LINEARSMOOTHER;
        _birrt.minimum_goal_paths = (int32_t)_parameters->_minimumgoalpaths;     // plannerparameters.h:740 (1 unless the caller's XML set it)
        _birrt.max_planning_time_ms = (int32_t)_parameters->_nMaxPlanningTime;    // milliseconds (planner.h; rrt.h:414-420)
        return PlannerStatus(PS_HasSolution);
    }

    virtual PlannerStatus PlanPath(TrajectoryBasePtr ptraj, int planningoptions = 0)
    {
        if (!_planner || !_parameters) return PlannerStatus("cudabirrt: not initialized", PS_Failed);
        const int ndof = _parameters->GetDOF();
        const int ninitial = (int)_parameters->vinitialconfig.size() / ndof, ngoals = (int)_parameters->vgoalconfig.size() / ndof;
        if (ninitial < 1 || (ngoals < 1 && !_parameters->_samplegoalfn) || (int)_parameters->vinitialconfig.size() % ndof ||
            (int)_parameters->vgoalconfig.size() % ndof)
            return PlannerStatus("cudabirrt: initial / goal configurations are improperly specified", PS_Failed);      // rrt.h:65-69, 322-328, 349-355
        // _samplegoalfn / _sampleinitialfn (planner.h:298-311): asked once per iteration, like BirrtPlanner::PlanPath does (rrt.h:422-437)
        cc_planner_set_samplers(_planner, !!_parameters->_samplegoalfn ? &CudaBirrtPlanner::_SampleGoalThunk : NULL,
                                !!_parameters->_sampleinitialfn ? &CudaBirrtPlanner::_SampleInitialThunk : NULL, this);
        cc_planner_set_progress_callback(_planner, &CudaBirrtPlanner::_ProgressThunk, this);       // RegisterPlanCallback (planner.h:700-716)
        // several initial configurations and several goals (one per IK solution, say) are one problem (rrt.h:72-88, 327-343)
        EnvironmentLock lock(GetEnv()->GetMutex());
        const int32_t start_begin[2] = {0, ninitial}, goal_begin[2] = {0, ngoals};
        const uint32_t seed = _parameters->_nRandomGeneratorSeed;
        cc_plan_result result;
        if (cc_birrt_plan_batch_multi(_planner, &_birrt, 1, start_begin, _parameters->vinitialconfig.data(), goal_begin,
                                      _parameters->vgoalconfig.data(), &seed, &result) != CC_OK)
            throw OPENRAVE_EXCEPTION_FORMAT("cudabirrt: %s", cc_last_error(), ORE_Failed);
        if (result.status == CC_PLAN_INTERRUPTED) return PlannerStatus("cudabirrt: planning was interrupted", PS_Interrupted);     // rrt.h:405-407
        if (result.status != CC_PLAN_SOLVED)
            return PlannerStatus("cudabirrt: no path", result.status == CC_PLAN_BAD_INITIAL ? (PS_Failed | PS_FailedDueToInitial)
                                                       : result.status == CC_PLAN_BAD_GOAL ? (PS_Failed | PS_FailedDueToGoal) : PS_Failed);
        int nw = 0;
        cc_planner_get_path(_planner, 0, 0, NULL, 0, &nw);
        std::vector<dReal> waypoints((size_t)nw * ndof);
        cc_planner_get_path(_planner, 0, 0, waypoints.data(), nw, &nw);
        if (ptraj->GetConfigurationSpecification().GetDOF() == 0) ptraj->Init(_parameters->_configurationspecification);
        ptraj->Insert(ptraj->GetNumWaypoints(), waypoints, _parameters->_configurationspecification);
        return PlannerStatus(PS_HasSolution);
    }

    /// starts / goals: [n][ndof]; every solved problem appends its waypoints to vtraj[k] (positions only; retiming is left to the
    /// caller's post-processor, as after linearsmoother with a retimer-less configuration)
    void PlanPathBatch(const std::vector<dReal>& starts, const std::vector<dReal>& goals, const std::vector<uint32_t>& seeds,
                       std::vector<TrajectoryBasePtr>& vtraj, std::vector<int>& vstatus)
    {
        EnvironmentLock lock(GetEnv()->GetMutex());
        const int ndof = _parameters->GetDOF();
        const int n = (int)starts.size() / ndof;
        std::vector<cc_plan_result> results(std::max(n, 1));
        if (cc_birrt_plan_batch(_planner, &_birrt, n, starts.data(), goals.data(), seeds.data(), results.data()) != CC_OK)
            throw OPENRAVE_EXCEPTION_FORMAT("cudabirrt: %s", cc_last_error(), ORE_Failed);
        vstatus.resize(n);
        std::vector<dReal> waypoints;
        for (int k = 0; k < n; ++k) {
            vstatus[k] = results[k].status;
            if (results[k].status != CC_PLAN_SOLVED || k >= (int)vtraj.size() || !vtraj[k]) continue;
            int nw = 0;
            cc_planner_get_path(_planner, k, 0, NULL, 0, &nw);
            waypoints.resize((size_t)nw * ndof);
            cc_planner_get_path(_planner, k, 0, waypoints.data(), nw, &nw);
            if (vtraj[k]->GetConfigurationSpecification().GetDOF() == 0) vtraj[k]->Init(_parameters->_configurationspecification);
            vtraj[k]->Insert(vtraj[k]->GetNumWaypoints(), waypoints, _parameters->_configurationspecification);
        }
    }

    virtual PlannerParametersConstPtr GetParameters() const { return _parameters; }

private:
    static int _SampleThunk(const boost::function<bool(std::vector<dReal>&)>& fn, int ndof, double* q_out)
    {
        std::vector<dReal> v;
        if (!fn(v) || (int)v.size() < ndof) return 0;
        std::copy(v.begin(), v.begin() + ndof, q_out);
        return 1;
    }
    static int _ProgressThunk(void* user, int, int iteration)
    {
        PlannerProgress progress;
        progress._iteration = iteration;
        return (int)static_cast<CudaBirrtPlanner*>(user)->_CallCallbacks(progress);     // PA_None / PA_Interrupt / PA_ReturnWithAnySolution
    }
    static int _SampleGoalThunk(void* user, int, int ndof, double* q_out) { return _SampleThunk(static_cast<CudaBirrtPlanner*>(user)->_parameters->_samplegoalfn, ndof, q_out); }
    static int _SampleInitialThunk(void* user, int, int ndof, double* q_out) { return _SampleThunk(static_cast<CudaBirrtPlanner*>(user)->_parameters->_sampleinitialfn, ndof, q_out); }

    RobotBasePtr _robot;
    RRTParametersPtr _parameters;
    cc_planner* _planner;
    cc_birrt_params _birrt;
};

// ---------------------------------------------------------------------------------------------- plugin entry (plugin.h:30-157)
class CudaCollisionPlugin : public RavePlugin
{
public:
    CudaCollisionPlugin()
    {
        _interfaces[PT_CollisionChecker].push_back("cudacollision"); _interfaces[PT_CollisionChecker].push_back("cudacachechecker");
        _interfaces[PT_Planner].push_back("cudabirrt"); _interfaces[PT_Planner].push_back("cudabatched");
    }
    virtual ~CudaCollisionPlugin() {}
    virtual InterfaceBasePtr CreateInterface(InterfaceType type, const std::string& interfacename, std::istream& sinput, EnvironmentBasePtr penv)
    {
        if (type == PT_CollisionChecker && interfacename == "cudacollision") return boost::make_shared<CudaCollisionChecker>(penv, boost::ref(sinput));
        if (type == PT_CollisionChecker && interfacename == "cudacachechecker") return boost::make_shared<CudaCacheChecker>(penv, boost::ref(sinput));
        if (type == PT_Planner && interfacename == "cudabirrt") return boost::make_shared<CudaBirrtPlanner>(penv);
        if (type == PT_Planner && interfacename == "cudabatched") return boost::make_shared<CudaBatchedPlanner>(penv, boost::ref(sinput));
        return InterfaceBasePtr();
    }
    virtual const InterfaceMap& GetInterfaces() const { return _interfaces; }
    virtual const std::string& GetPluginName() const { static const std::string name("CudaCollisionPlugin"); return name; }

private:
    InterfaceMap _interfaces;
};

OPENRAVE_PLUGIN_API RavePlugin* CreatePlugin() { return new CudaCollisionPlugin(); }
