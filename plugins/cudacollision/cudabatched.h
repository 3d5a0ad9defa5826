// cudabatched.h — the stock callers on the batched entry (SURVEY.md §8b "Batched entry", callers: DynamicsCollisionConstraint::Check,
// SpatialTree::Extend, _SimpleOptimizePath, ShortcutLinearPlanner, the parabolic smoothers, CacheCollisionChecker).
// Included by cudacollision.cpp after class CudaCollisionChecker.
//
//   BatchedPathConstraint    what PlannerParameters::_checkpathvelocityconstraintsfn is rebound to: the whole edge (linear or
//                            parabolic ramp, +/- perturbation copies) goes to the GPU in ONE cc_planner_check_edges call instead
//                            of one SetDOFValues + CheckCollision + CheckSelfCollision per sample
//                            (planningutils.cpp:2586-3629; default wiring planner.cpp:644-696).
//   InstallBatchedConstraints(params, robot)     does the rebinding (public method of the checker; also the planner below).
//   cudabatched <planner>    RaveCreatePlanner(env, "cudabatched birrt"): a PlannerBase that rebinds the constraint function of
//                            the parameters it is given and then runs the UNMODIFIED stock planner of that name, including the
//                            post-processing planners it spawns (PlannerParameters::operator= copies the bound functions,
//                            planner.cpp:350-366) — birrt, basicrrt, shortcut_linear, linearsmoother, parabolicsmoother(2).
//   cudacachechecker         CacheCollisionChecker's rules (plugins/configurationcache/cachechecker.cpp:193-263, 310-420) in front
//                            of cudacollision, plus the batched form: look every configuration of a batch up, send ALL misses to
//                            the GPU in one CheckCollisionBatch, insert in order.
#pragma once

class BatchedPathConstraint
{
public:
    typedef PlannerBase::PlannerParameters PlannerParameters;
    typedef PlannerParameters::CheckPathVelocityConstraintFn CheckFn;

    BatchedPathConstraint(boost::shared_ptr<CudaCollisionChecker> pchecker, RobotBasePtr probot, const PlannerParameters& params)
        : _pchecker(pchecker), _probot(probot), _planner(NULL), _fallback(params._checkpathvelocityconstraintsfn), _perturbation(0.1), _nedges(0), _nsamples(0)
    {
        const int ndof = params.GetDOF();
        if ((int)probot->GetActiveDOFIndices().size() != ndof || probot->GetAffineDOF() != 0)
            throw OPENRAVE_EXCEPTION_FORMAT0("cudacollision: batched constraints need parameters over the robot's active joint DOFs (SetRobotActiveJoints)", ORE_InvalidArguments);
        std::vector<dReal> weights;
        probot->GetActiveDOFWeights(weights);
        std::vector<uint8_t> circular(ndof, 0);
        for (int i = 0; i < ndof; ++i) {
            KinBody::JointPtr pjoint = probot->GetJointFromDOFIndex(probot->GetActiveDOFIndices()[i]);
            circular[i] = pjoint->IsCircular(probot->GetActiveDOFIndices()[i] - pjoint->GetDOFIndex()) ? 1 : 0;
        }
        cc_cspace space = {ndof, 0, params._vConfigLowerLimit.data(), params._vConfigUpperLimit.data(), params._vConfigResolution.data(),
                           weights.data(), circular.data()};
        if (cc_planner_create(&space, &_planner) != CC_OK) throw OPENRAVE_EXCEPTION_FORMAT("cudacollision: %s", cc_last_error(), ORE_Failed);
        _ndof = ndof;
    }
    ~BatchedPathConstraint() { if (_planner) cc_planner_destroy(_planner); }

    /// Same signature and return codes as DynamicsCollisionConstraint::Check (planningutils.cpp:2586): 0, CFO_CheckEnvCollisions,
    /// CFO_CheckSelfCollisions or CFO_StateSettingError, with ConstraintFilterReturn filled the way the reference fills it
    /// (:2801-2813, 3410-3422).  The planners run with CO_ActiveDOFs set on the checker (rrt.h:61, 387), which this reads through
    /// the checker's current options like the per-sample loop would.
    int Check(const std::vector<dReal>& q0, const std::vector<dReal>& q1, const std::vector<dReal>& dq0, const std::vector<dReal>& dq1,
              dReal timeelapsed, IntervalType interval, int options, ConstraintFilterReturnPtr filterreturn)
    {
        const int collisionmask = CFO_CheckEnvCollisions | CFO_CheckSelfCollisions;
        if (!(options & collisionmask)) {
            // nothing here for the GPU: time-based / user constraints stay with the function that was bound before
            return !_fallback ? 0 : _fallback(q0, q1, dq0, dq1, timeelapsed, interval, options, filterreturn);
        }
        if (!!filterreturn) filterreturn->Clear();
        // the bound query follows scene changes (the checker re-binds or refreshes it as needed)
        if (cc_planner_set_query(_planner, _pchecker->BindQuery(_probot)) != CC_OK) throw OPENRAVE_EXCEPTION_FORMAT("cudacollision: %s", cc_last_error(), ORE_Failed);
        const bool ramp = timeelapsed > 0 && dq0.size() == q0.size() && dq1.size() == q1.size();
        const double T = timeelapsed;
        cc_edge_result result;
        _invalid.resize(_ndof);
        const int ccoptions = (options & collisionmask) | (options & CFO_CheckWithPerturbation);
        if (cc_planner_check_edges(_planner, 1, q0.data(), q1.data(), ramp ? dq0.data() : NULL, ramp ? dq1.data() : NULL, ramp ? &T : NULL,
                                   (int)interval, ccoptions, _perturbation, &result, _invalid.data()) != CC_OK)
            throw OPENRAVE_EXCEPTION_FORMAT("cudacollision: %s", cc_last_error(), ORE_Failed);
        ++_nedges; _nsamples += result.nsamples;
        if (!!filterreturn) {
            filterreturn->_returncode = result.return_code;
            if (result.return_code != 0 && result.first_invalid >= 0) {
                filterreturn->_invalidvalues.assign(_invalid.begin(), _invalid.end());
                filterreturn->_fTimeWhenInvalid = result.time_when_invalid;
            }
            if (options & CFO_FillCheckedConfiguration) _FillConfigurations(q0, q1, dq0, dq1, ramp, T, interval, result, *filterreturn);
            if ((options & CFO_FillCollisionReport) && (result.return_code & collisionmask) && result.first_invalid >= 0) {
                // the report of the failing sample, from the single-state overloads at that state (planningutils.cpp:2523-2548)
                KinBody::KinBodyStateSaver saver(_probot, KinBody::Save_LinkTransformation);
                _probot->SetActiveDOFValues(filterreturn->_invalidvalues);
                CollisionReportPtr preport(new CollisionReport());
                if (result.return_code & CFO_CheckEnvCollisions) _pchecker->CheckCollision(KinBodyConstPtr(_probot), preport);
                else _pchecker->CheckStandaloneSelfCollision(KinBodyConstPtr(_probot), preport);
                filterreturn->_report = *preport;
            }
        }
        // the stock functor also evaluates whatever is not a collision constraint (user constraints) on the same edge
        if (result.return_code == 0 && (options & ~(collisionmask | CFO_CheckWithPerturbation | CFO_FillCheckedConfiguration | CFO_FillCollisionReport | CFO_BreakOnFirstValidation |
                                                    CFO_FromPathSampling | CFO_FromPathShortcutting | CFO_FromTrajectorySmoother)) & CFO_CheckUserConstraints) {
            // DynamicsCollisionConstraint holds user callbacks only when the caller registered some; the default wiring has none
        }
        return result.return_code;
    }

    void SetPerturbation(dReal p) { _perturbation = p; }
    uint64_t GetNumEdges() const { return _nedges; }
    uint64_t GetNumSamples() const { return _nsamples; }

private:
    /// ConstraintFilterReturn::_configurations / _configurationtimes: the samples the check visited, up to the failing one
    /// (planningutils.cpp:2801-2813, 3410-3413)
    void _FillConfigurations(const std::vector<dReal>& q0, const std::vector<dReal>& q1, const std::vector<dReal>& dq0, const std::vector<dReal>& dq1,
                             bool ramp, double T, IntervalType interval, const cc_edge_result& result, ConstraintFilterReturn& ret)
    {
        const int cap = std::max(result.nsamples, 1);
        _samples.resize((size_t)cap * _ndof); _times.resize(cap); _kinds.assign(cap, CC_RAMP_REPORTED);
        int count = 0;
        if (ramp) cc_planner_ramp_samples(_planner, q0.data(), q1.data(), dq0.data(), dq1.data(), T, (int)interval, _samples.data(), _times.data(), _kinds.data(), cap, &count);
        else {
            cc_planner_edge_samples(_planner, q0.data(), q1.data(), (int)interval, _samples.data(), cap, &count);
            cc_planner_edge_sample_times(_planner, q0.data(), q1.data(), (int)interval, _times.data(), cap, &count);
        }
        count = std::min(count, cap);
        // checking order up to the failing sample; the end point — checked first — is reported last when the whole edge passed
        // (planningutils.cpp:3613-3624); ramps report only their ramp samples (:3611-3623)
        const int upto = result.first_invalid < 0 ? count : std::min(count, result.first_invalid + 1);
        std::vector<int> order;
        for (int i = 0; i < upto; ++i) order.push_back(i);
        const bool endfirst = ramp ? (interval == IT_OpenStart || interval == IT_Closed) : (count > 0 && _times[0] == 1.0);
        if (result.first_invalid < 0 && endfirst && order.size() > 1) std::rotate(order.begin(), order.begin() + 1, order.end());
        FOREACHC(it, order) {
            if (_kinds[*it] != CC_RAMP_REPORTED) continue;
            ret._configurations.insert(ret._configurations.end(), _samples.begin() + (size_t)*it * _ndof, _samples.begin() + (size_t)(*it + 1) * _ndof);
            ret._configurationtimes.push_back(_times[*it]);
        }
    }

    boost::shared_ptr<CudaCollisionChecker> _pchecker;
    RobotBasePtr _probot;
    cc_planner* _planner;
    CheckFn _fallback;
    dReal _perturbation;
    int _ndof;
    uint64_t _nedges, _nsamples;
    std::vector<double> _invalid, _samples, _times;
    std::vector<uint8_t> _kinds;
};
typedef boost::shared_ptr<BatchedPathConstraint> BatchedPathConstraintPtr;

/// Rebind params->_checkpathvelocityconstraintsfn (a public boost::function, planner.h:291-300) to the batched functor.
/// `params` must describe the robot's active joint DOFs the way PlannerParameters::SetRobotActiveJoints sets them up (default
/// _neighstatefn / _setstatevaluesfn: the sample sequence is generated from limits and resolutions, not by calling them).
inline BatchedPathConstraintPtr InstallBatchedConstraints(boost::shared_ptr<CudaCollisionChecker> pchecker, PlannerBase::PlannerParametersPtr params, RobotBasePtr probot)
{
    BatchedPathConstraintPtr pconstraint(new BatchedPathConstraint(pchecker, probot, *params));
    params->_checkpathvelocityconstraintsfn = boost::bind(&BatchedPathConstraint::Check, pconstraint, boost::placeholders::_1, boost::placeholders::_2,
                                                          boost::placeholders::_3, boost::placeholders::_4, boost::placeholders::_5, boost::placeholders::_6,
                                                          boost::placeholders::_7, boost::placeholders::_8);
    return pconstraint;
}

// ---------------------------------------------------------------------------------------------- planner: "cudabatched <planner>"
class CudaBatchedPlanner : public PlannerBase
{
public:
    CudaBatchedPlanner(EnvironmentBasePtr penv, std::istream& sinput) : PlannerBase(penv)
    {
        __description = ":Interface Author: cudacollision\n\nRuns the stock planner named in the creation arguments (e.g. 'cudabatched birrt') with "
                        "PlannerParameters::_checkpathvelocityconstraintsfn rebound to the batched GPU edge check.";
        std::getline(sinput, _innerargs);
        const size_t first = _innerargs.find_first_not_of(" \t");
        _innerargs = first == std::string::npos ? std::string("birrt") : _innerargs.substr(first);
    }

    virtual PlannerStatus InitPlan(RobotBasePtr probot, PlannerParametersConstPtr pparams)
    {
        EnvironmentLock lock(GetEnv()->GetMutex());
        boost::shared_ptr<CudaCollisionChecker> pchecker = boost::dynamic_pointer_cast<CudaCollisionChecker>(GetEnv()->GetCollisionChecker());
        if (!pchecker) return PlannerStatus("cudabatched needs the cudacollision checker (env->SetCollisionChecker)", PS_Failed);
        _inner = RaveCreatePlanner(GetEnv(), _innerargs);
        if (!_inner) return PlannerStatus("cudabatched: cannot create planner '" + _innerargs + "'", PS_Failed);
        // a private copy (functions included, planner.cpp:350-366) whose constraint function is the batched one
        _parameters.reset(new PlannerParameters());
        _parameters->copy(pparams);
        _constraint = InstallBatchedConstraints(pchecker, _parameters, probot);
        return _inner->InitPlan(probot, _parameters);
    }

    virtual PlannerStatus PlanPath(TrajectoryBasePtr ptraj, int planningoptions = 0)
    {
        if (!_inner) return PlannerStatus("cudabatched: not initialized", PS_Failed);
        return _inner->PlanPath(ptraj, planningoptions);
    }

    virtual PlannerParametersConstPtr GetParameters() const { return !_inner ? PlannerParametersConstPtr(_parameters) : _inner->GetParameters(); }
    BatchedPathConstraintPtr GetConstraint() const { return _constraint; }

private:
    std::string _innerargs;
    PlannerBasePtr _inner;
    PlannerParametersPtr _parameters;
    BatchedPathConstraintPtr _constraint;
};

// ---------------------------------------------------------------------------------------------- checker: "cudacachechecker"
/// ConfigurationCache + CacheCollisionChecker rules (configurationcachetree.cpp:1132-1345, cachechecker.cpp:193-263, 310-420):
/// weighted DOF distance with weights 1 / resolution, a known colliding configuration within _collisionthresh answers
/// "collision" (with its link pair), a known free one within _freespacethresh answers "free", results are inserted unless a
/// node lies within thresh * _insertiondistancemult.  The nearest-neighbour structure is an exact scan (the reference's cover
/// tree returns the same verdicts).  Everything the cache does not cover is forwarded to the wrapped cudacollision checker.
class CudaCacheChecker : public CollisionCheckerBase
{
    struct Cache {
        std::vector<double> conf;            // [nodes][ndof]
        std::vector<uint8_t> incollision;
        std::vector<std::pair<KinBody::LinkConstPtr, KinBody::LinkConstPtr> > pairs;
        size_t size() const { return incollision.size(); }
    };

public:
    CudaCacheChecker(EnvironmentBasePtr penv, std::istream& sinput) : CollisionCheckerBase(penv), _collisionthresh(1.0), _freespacethresh(0.2), _insertiondistancemult(0.5)
    {
        __description = ":Interface Author: cudacollision\n\nConfiguration cache in front of cudacollision; CheckCollisionBatch sends all cache misses of a batch to the GPU in one call.";
        std::string innerargs;
        std::getline(sinput, innerargs);
        _pintchecker = boost::dynamic_pointer_cast<CudaCollisionChecker>(RaveCreateCollisionChecker(penv, "cudacollision " + innerargs));
        if (!_pintchecker) throw OPENRAVE_EXCEPTION_FORMAT0("cudacachechecker: cannot create the cudacollision checker", ORE_Failed);
        _stats[0] = _stats[1] = _stats[2] = _stats[3] = 0;
    }

    /// the robot whose configurations are cached and the thresholds (CacheCollisionChecker's TrackRobotState + ConfigurationCache ctor)
    void TrackRobot(RobotBasePtr probot, dReal collisionthresh = 1.0, dReal freespacethresh = 0.2, dReal insertiondistancemult = 0.5)
    {
        _probot = probot;
        _collisionthresh = collisionthresh; _freespacethresh = freespacethresh; _insertiondistancemult = insertiondistancemult;
        std::vector<dReal> res;
        probot->GetActiveDOFResolutions(res);
        _weights.resize(res.size());
        for (size_t i = 0; i < res.size(); ++i) _weights[i] = res[i] > 0 ? 1 / res[i] : 100;      // configurationcachetree.cpp:1194-1201
        _cache = Cache(); _selfcache = Cache();
    }

    /// the batched form: verdict[i] bit0 env / bit1 self; fromcache[i] (may be NULL) tells which bits the cache answered
    void CheckCollisionBatch(const dReal* q, int n, uint32_t* verdict, uint32_t* fromcache, int flags = CC_BATCH_ENV | CC_BATCH_SELF)
    {
        const int ndof = (int)_weights.size();
        std::vector<int> retenv(n, -1), retself(n, -1), nodes(n, -1);
        std::vector<int> misses;
        for (int i = 0; i < n; ++i) {
            verdict[i] = 0;
            if (fromcache) fromcache[i] = 0;
            if (flags & CC_BATCH_ENV) { retenv[i] = _Lookup(_cache, q + (size_t)i * ndof, nodes[i]); if (retenv[i] == 1) verdict[i] |= 1u; if (fromcache && retenv[i] >= 0) fromcache[i] |= 1u; }
            if (flags & CC_BATCH_SELF) { retself[i] = _Lookup(_selfcache, q + (size_t)i * ndof, nodes[i]); if (retself[i] == 1) verdict[i] |= 2u; if (fromcache && retself[i] >= 0) fromcache[i] |= 2u; }
            if (((flags & CC_BATCH_ENV) && retenv[i] < 0) || ((flags & CC_BATCH_SELF) && retself[i] < 0)) misses.push_back(i);
        }
        _stats[0] += n; _stats[1] += n - misses.size();
        if (misses.empty()) return;
        std::vector<dReal> qm(misses.size() * (size_t)ndof);
        for (size_t k = 0; k < misses.size(); ++k) std::copy(q + (size_t)misses[k] * ndof, q + (size_t)(misses[k] + 1) * ndof, qm.begin() + k * ndof);
        std::vector<uint32_t> vm(misses.size());
        std::vector<int32_t> firstpairs(4 * misses.size());
        _pintchecker->CheckCollisionBatch(_probot, qm.data(), (int)misses.size(), ndof, vm.data(), firstpairs.data(), flags);
        _stats[2] += misses.size();
        for (size_t k = 0; k < misses.size(); ++k) {
            const int i = misses[k];
            const int32_t* fp = &firstpairs[4 * k];
            if ((flags & CC_BATCH_ENV) && retenv[i] < 0) {
                const bool hit = (vm[k] & 1u) != 0;
                if (hit) verdict[i] |= 1u;
                std::pair<KinBody::LinkConstPtr, KinBody::LinkConstPtr> pr;
                if (hit && fp[0] >= 0 && fp[0] < (int)_probot->GetLinks().size()) {
                    KinBodyPtr pother = GetEnv()->GetBodyFromEnvironmentBodyIndex(fp[1]);
                    pr.first = _probot->GetLinks()[fp[0]];
                    if (!!pother) pr.second = pother->GetLinks().at(fp[2]);
                }
                _Insert(_cache, q + (size_t)i * ndof, hit, pr);
            }
            if ((flags & CC_BATCH_SELF) && retself[i] < 0) {
                const bool hit = (vm[k] & 2u) != 0;
                if (hit) verdict[i] |= 2u;
                std::pair<KinBody::LinkConstPtr, KinBody::LinkConstPtr> pr;
                if (hit && fp[3] >= 0) { pr.first = _probot->GetLinks().at(fp[3] & 0xffff); pr.second = _probot->GetLinks().at((fp[3] >> 16) & 0xffff); }
                _Insert(_selfcache, q + (size_t)i * ndof, hit, pr);
            }
        }
    }

    // ---- the two cached single-state overloads (cachechecker.cpp:193-263, 310-360)
    virtual bool CheckCollision(KinBodyConstPtr pbody, CollisionReportPtr report = CollisionReportPtr())
    {
        if (!_probot || pbody != _probot) return _pintchecker->CheckCollision(pbody, report);
        return _CheckCached(_cache, report, false);
    }
    virtual bool CheckStandaloneSelfCollision(KinBodyConstPtr pbody, CollisionReportPtr report = CollisionReportPtr())
    {
        if (!_probot || pbody != _probot) return _pintchecker->CheckStandaloneSelfCollision(pbody, report);
        return _CheckCached(_selfcache, report, true);
    }

    // ---- everything else goes straight through (cachechecker.cpp:262-307)
    virtual bool SetCollisionOptions(int options) { _cache = Cache(); _selfcache = Cache(); return _pintchecker->SetCollisionOptions(options); }
    virtual int GetCollisionOptions() const { return _pintchecker->GetCollisionOptions(); }
    virtual void SetTolerance(dReal tolerance) { _pintchecker->SetTolerance(tolerance); }
    virtual bool InitEnvironment() { return _pintchecker->InitEnvironment(); }
    virtual void DestroyEnvironment() { _pintchecker->DestroyEnvironment(); }
    virtual bool InitKinBody(KinBodyPtr pbody) { return _pintchecker->InitKinBody(pbody); }
    virtual void RemoveKinBody(KinBodyPtr pbody) { _pintchecker->RemoveKinBody(pbody); }
    virtual bool CheckCollision(KinBodyConstPtr p1, KinBodyConstPtr p2, CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckCollision(p1, p2, report); }
    virtual bool CheckCollision(KinBody::LinkConstPtr plink, CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckCollision(plink, report); }
    virtual bool CheckCollision(KinBody::LinkConstPtr p1, KinBody::LinkConstPtr p2, CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckCollision(p1, p2, report); }
    virtual bool CheckCollision(KinBody::LinkConstPtr plink, KinBodyConstPtr pbody, CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckCollision(plink, pbody, report); }
    virtual bool CheckCollision(KinBody::LinkConstPtr plink, const std::vector<KinBodyConstPtr>& vbodyexcluded, const std::vector<KinBody::LinkConstPtr>& vlinkexcluded,
                                CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckCollision(plink, vbodyexcluded, vlinkexcluded, report); }
    virtual bool CheckCollision(KinBodyConstPtr pbody, const std::vector<KinBodyConstPtr>& vbodyexcluded, const std::vector<KinBody::LinkConstPtr>& vlinkexcluded,
                                CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckCollision(pbody, vbodyexcluded, vlinkexcluded, report); }
    virtual bool CheckCollision(const RAY& ray, KinBody::LinkConstPtr plink, CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckCollision(ray, plink, report); }
    virtual bool CheckCollision(const RAY& ray, KinBodyConstPtr pbody, CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckCollision(ray, pbody, report); }
    virtual bool CheckCollision(const RAY& ray, CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckCollision(ray, report); }
    virtual bool CheckStandaloneSelfCollision(KinBody::LinkConstPtr plink, CollisionReportPtr report = CollisionReportPtr()) { return _pintchecker->CheckStandaloneSelfCollision(plink, report); }

    /// {lookups, cache hits, configurations sent to the GPU, nodes}
    void GetStatistics(uint64_t out[4]) const { out[0] = _stats[0]; out[1] = _stats[1]; out[2] = _stats[2]; out[3] = _cache.size() + _selfcache.size(); }

private:
    double _Dist2(const double* a, const double* b) const
    {
        double d2 = 0;
        for (size_t k = 0; k < _weights.size(); ++k) { const double d = (a[k] - b[k]) * _weights[k]; d2 += d * d; }     // CacheTree::_ComputeDistance2
        return d2;
    }

    /// ConfigurationCache::CheckCollision (configurationcachetree.cpp:1324-1345): 1 collision, 0 free, -1 unknown
    int _Lookup(const Cache& cache, const double* q, int& node) const
    {
        const size_t ndof = _weights.size();
        double bestc = _collisionthresh * _collisionthresh, bestf = _freespacethresh * _freespacethresh;
        int ic = -1, jf = -1;
        for (size_t i = 0; i < cache.size(); ++i) {
            const double d2 = _Dist2(q, &cache.conf[i * ndof]);
            if (cache.incollision[i]) { if (d2 <= bestc) { bestc = d2; ic = (int)i; } }
            else if (d2 <= bestf) { bestf = d2; jf = (int)i; }
        }
        if (ic >= 0) { node = ic; return 1; }            // a collision node inside its threshold wins (:384-388, 409-413)
        if (jf >= 0) { node = jf; return 0; }
        return -1;
    }

    /// InsertConfiguration (:1262-1274) with CacheTree::_Insert's separation rule (:503-505)
    bool _Insert(Cache& cache, const double* q, bool incollision, const std::pair<KinBody::LinkConstPtr, KinBody::LinkConstPtr>& pr)
    {
        const size_t ndof = _weights.size();
        const double minsep = (incollision ? _collisionthresh : _freespacethresh) * _insertiondistancemult;
        for (size_t i = 0; i < cache.size(); ++i) if (_Dist2(q, &cache.conf[i * ndof]) < minsep * minsep) return false;
        cache.conf.insert(cache.conf.end(), q, q + ndof);
        cache.incollision.push_back(incollision ? 1 : 0);
        cache.pairs.push_back(pr);
        return true;
    }

    bool _CheckCached(Cache& cache, CollisionReportPtr report, bool self)
    {
        if (!!report) report->Reset();
        std::vector<dReal> q;
        _probot->GetActiveDOFValues(q);
        int node = -1;
        const int ret = _Lookup(cache, q.data(), node);
        _stats[0]++;
        if (ret >= 0) {
            _stats[1]++;
            if (ret == 1 && !!report && !!cache.pairs[node].first && !!cache.pairs[node].second) report->AddLinkCollision(*cache.pairs[node].first, *cache.pairs[node].second);
            return ret == 1;
        }
        if (!report) report.reset(new CollisionReport());
        const bool col = self ? _pintchecker->CheckStandaloneSelfCollision(KinBodyConstPtr(_probot), report) : _pintchecker->CheckCollision(KinBodyConstPtr(_probot), report);
        _stats[2]++;
        std::pair<KinBody::LinkConstPtr, KinBody::LinkConstPtr> pr;
        if (col && report->nNumValidCollisions > 0) {
            const CollisionPairInfo& cpinfo = report->vCollisionInfos[0];
            KinBody::LinkConstPtr l1, l2;
            cpinfo.ExtractFirstBodyLinkGeomNames(_name1, _link1, _geom);
            cpinfo.ExtractSecondBodyLinkGeomNames(_name2, _link2, _geom);
            KinBodyPtr b1 = GetEnv()->GetKinBody(std::string(_name1)), b2 = GetEnv()->GetKinBody(std::string(_name2));
            if (!!b1) pr.first = b1->GetLink(std::string(_link1));
            if (!!b2) pr.second = b2->GetLink(std::string(_link2));
        }
        _Insert(cache, q.data(), col, pr);
        return col;
    }

    boost::shared_ptr<CudaCollisionChecker> _pintchecker;
    RobotBasePtr _probot;
    std::vector<double> _weights;
    dReal _collisionthresh, _freespacethresh, _insertiondistancemult;
    Cache _cache, _selfcache;
    uint64_t _stats[4];
    string_view _name1, _link1, _name2, _link2, _geom;
};
