"""Edge discretisation: the batched replacement of `DynamicsCollisionConstraint::Check` (linear branch).

The reference checks an edge q0 -> q1 one sample at a time (src/libopenrave/planningutils.cpp:2586-3629):
  * `bCheckEnd` intervals test q1 FIRST (:2692-2708); then q0 when the interval is closed at the start (:2809-2826);
  * the number of steps is max over DOFs of int(|dq_i| / resolution_i + 0.99) (:2772-2786), dq from `_diffstatefn`
    (circular DOFs wrap, robot.cpp:1639-1662);
  * samples f = start .. numSteps-1 are produced by `_neighstatefn` = AddStatesWithLimitCheck (planner.cpp:88-112, 687) from
    the previous sample with the delta re-derived from q0 + (f+1) dQ (:3438-3452);
  * each sample runs `_SetAndCheckState` -> env collision, then self collision (:2317-2368, 2523-2548), optionally again at
    +/- perturbation * resolution (CFO_CheckWithPerturbation, :2344-2366), and the first failure returns.
With the default neighbour function the sample sequence does not depend on any collision result, so the whole edge (or
many edges) is generated up front and sent through `CheckCollisionBatch` in one call; the return code, the failing
sample and `_fTimeWhenInvalid` are then those of the first failing sample in the reference's order.
`Check` is that batched path; `CheckSerial` is the literal per-sample loop over any CollisionCheckerBase-shaped checker
and exists so the two can be compared (tier V6).
"""
import math

import numpy as np

IT_Open, IT_OpenStart, IT_OpenEnd, IT_Closed = 0, 1, 2, 3                 # planner.h IntervalType
CFO_CheckEnvCollisions, CFO_CheckSelfCollisions = 1, 2                    # planner.h:32-33
CFO_CheckWithPerturbation = 0x00010000                                    # planner.h:37
CFO_FillCheckedConfiguration = 0x00020000                                 # planner.h:38
CFO_StateSettingError = 0x80000000                                        # planner.h:44
G_EPSILON_JOINT_LIMIT = 1e-15 ** 0.8                                       # g_fEpsilonJointLimit (libopenrave.h:141)
G_EPSILON_LINEAR = 1e-15 ** 0.9                                            # g_fEpsilonLinear (libopenrave.h:140)


def _solvequad(a, b, c):
    """mathextra::solvequad (include/openrave/mathextra.h:214-256); returns the list of roots."""
    if a == 0:
        return [] if b == 0 else [-c / b]
    if c == 0:
        return [0.0, -b / a]
    d = b * b - 4.0 * c * a
    if d < 0:
        if d < -1e-16:
            return []
        d = 0.0
    d = math.sqrt(d)
    r1 = 0.5 * (-b + d) / a if abs(-b - d) < abs(a) else 2.0 * c / (-b - d)
    r2 = 0.5 * (-b - d) / a if abs(-b + d) < abs(a) else 2.0 * c / (-b + d)
    return [r1, r2]


class ConstraintFilterReturn(object):
    """planner.h:102-130."""

    def __init__(self):
        self.Clear()

    def Clear(self):
        self._configurations = []
        self._configurationtimes = []
        self._invalidvalues = []
        self._fTimeWhenInvalid = 0.0
        self._returncode = 0


class DynamicsCollisionConstraint(object):
    """planningutils.h DynamicsCollisionConstraint for one robot over its active DOFs (PlannerParameters::SetRobotActiveJoints,
    planner.cpp:636-697): limits, resolutions and the circular mask come from the robot."""

    def __init__(self, robot, checker, filtermask=CFO_CheckEnvCollisions | CFO_CheckSelfCollisions | CFO_CheckWithPerturbation |
                 CFO_FillCheckedConfiguration, perturbation=0.1):
        self.robot = robot
        self.checker = checker
        self.filtermask = filtermask
        self.perturbation = perturbation                                   # planningutils.cpp:2298 default
        self.active = robot.GetActiveDOFIndices()
        joints = [robot.joints[d] for d in self.active]
        self.lower = np.array([j.lower for j in joints])
        self.upper = np.array([j.upper for j in joints])
        self.resolution = np.array([j.resolution for j in joints])
        self.circular = np.array([bool(j.circular) for j in joints])
        self.weights = np.array([j.weight for j in joints])

    # ---- parameter functions
    def diffstate(self, q1, q0):
        d = np.array(q1, dtype=np.float64) - np.array(q0, dtype=np.float64)
        if self.circular.any():                                            # Joint::SubtractValue -> NormalizeCircularAngle
            w = d[self.circular]
            w = np.where(w < -math.pi, w + 2 * math.pi * np.ceil((-math.pi - w) / (2 * math.pi)), w)
            w = np.where(w >= math.pi, w - 2 * math.pi * np.floor((w + math.pi) / (2 * math.pi)), w)
            d[self.circular] = w
        return d

    def neighstate(self, q, delta):
        """AddStatesWithLimitCheck (planner.cpp:88-112): returns (q', deviated)."""
        q = q + delta
        dev = bool(((q > self.upper + G_EPSILON_JOINT_LIMIT) | (q < self.lower - G_EPSILON_JOINT_LIMIT)).any())
        return np.minimum(np.maximum(q, self.lower), self.upper), dev

    def num_steps(self, dQ):
        steps = np.where(self.resolution != 0, (np.abs(dQ) / np.where(self.resolution != 0, self.resolution, 1.0) + 0.99).astype(np.int64),
                         (np.abs(dQ) * 100).astype(np.int64))
        return int(steps.max()) if len(steps) else 0, int(steps.sum())

    # ---- the sample sequence of one edge, in the order the reference checks it
    def edge_samples(self, q0, q1, interval=IT_Closed):
        """Returns (samples [m, ndof], times [m], kinds [m]) with kinds 'end' | 'start' | 'step'."""
        q0 = np.array(q0, dtype=np.float64)
        q1 = np.array(q1, dtype=np.float64)
        start = 1 if interval in (IT_Open, IT_OpenStart) else 0
        check_end = interval in (IT_OpenStart, IT_Closed)
        out, times, kinds = [], [], []
        if check_end:
            out.append(q1.copy()); times.append(1.0); kinds.append('end')
        dQ = self.diffstate(q1, q0)
        nsteps, total = self.num_steps(dQ)
        if total == 0 and start > 0:
            return np.array(out).reshape(-1, len(q0)), times, kinds
        if start == 0:
            out.append(q0.copy()); times.append(0.0); kinds.append('start')
            start = 1
        if nsteps == 0:
            return np.array(out).reshape(-1, len(q0)), times, kinds
        fisteps = 1.0 / nsteps
        dQ = dQ * fisteps
        cur, dev = self.neighstate(q0.copy(), dQ)
        if dev:
            raise NotImplementedError("edge leaves the joint limits (NSS_SuccessfulWithDeviation sub-stepping, planningutils.cpp:3454-3519)")
        for f in range(start, nsteps):
            out.append(cur.copy()); times.append(f * fisteps); kinds.append('step')
            prev = q0 + (f + 1) * dQ - cur
            big = np.abs(prev) > self.resolution
            if big.any():
                prev = prev * float((self.resolution[big] / np.abs(prev[big])).min())
            cur, dev = self.neighstate(cur, prev)
            if dev:
                raise NotImplementedError("edge leaves the joint limits (NSS_SuccessfulWithDeviation sub-stepping)")
        if float(np.sqrt(((cur - q1) ** 2).sum())) > 1e-7:                  # :3524-3541 (SimpleDistanceMetric with unit weights)
            out.append(cur.copy()); times.append(1.0); kinds.append('step')
        return np.array(out).reshape(-1, len(q0)), times, kinds

    # ---- the sample sequence of a parabolic ramp (quadratic branch), in the order the reference checks it
    def ramp_samples(self, q0, q1, dq0, dq1, timeelapsed, interval=IT_Closed):
        """DynamicsCollisionConstraint::Check with end-point velocities and timeelapsed > 0 (planningutils.cpp:2660-2681,
        2726-2766, 2844-3319).  Returns (samples [m, ndof], times [m] (NaN = linear sub-step), kinds [m]: 1 reported, 0 sub-step, 2 tail sub-step,
        3 unreported end point) or None for
        CFO_StateSettingError.  Scalar arithmetic in the reference's order (csrc/cc_planner.cpp holds the C++ twin)."""
        n = len(self.active)
        q0 = [float(x) for x in q0]; q1 = [float(x) for x in q1]
        dq0 = [float(x) for x in dq0]; dq1 = [float(x) for x in dq1]
        T = float(timeelapsed)
        res = [float(x) for x in self.resolution]
        eps_lin = G_EPSILON_LINEAR
        start = 1 if interval in (IT_Open, IT_OpenStart) else 0
        check_end = interval in (IT_OpenStart, IT_Closed)
        out, times, filled = [], [], []

        def push(q, t, f):
            out.append(list(q)); times.append(t); filled.append(int(f))

        def pack():
            return np.array(out, dtype=np.float64).reshape(-1, n), times, filled

        if check_end:
            push(q1, T, 1)
        dQ = [float(x) for x in self.diffstate(q1, q0)]
        iT = 1.0 / T
        acc = [(dq1[i] - dq0[i]) * iT for i in range(n)]
        nsteps, largest, total = 0, -1, 0
        for i in range(n):
            travel = abs(dQ[i])
            if abs(acc[i]) > eps_lin:
                tin = -dq0[i] / acc[i]
                if 0 <= tin < T:
                    pin = 0.5 * dq0[i] * tin
                    travel = abs(pin) + abs(dQ[i] - pin)
            st = int(travel / res[i] + 0.99) if res[i] != 0 else int(travel * 100)
            total += st
            if st > nsteps:
                nsteps, largest = st, i
        if total == 0 and start > 0:
            return pack()
        if start == 0:
            push(q0, 0.0, 1)
            start = 1
        if nsteps == 0:
            return pack()
        aL, vL, largest_step = acc[largest], dq0[largest], dQ[largest]
        t_infl, f_infl = T, dQ[largest]
        travel = f_infl
        if abs(aL) > eps_lin:
            tin = -vL / aL
            if 0 <= tin < T:
                t_infl = tin
                f_infl = 0.5 * vL * t_infl
                travel = abs(f_infl) + abs(dQ[largest] - f_infl)
                fdelta = travel / nsteps
                if f_infl < 0 or (abs(f_infl) <= eps_lin and dQ[largest] > 0):
                    fdelta = -fdelta
            else:
                fdelta = travel / float(nsteps)
        else:
            fdelta = travel / float(nsteps)
        cur, vel = list(q0), list(dq0)
        lower, upper = [float(x) for x in self.lower], [float(x) for x in self.upper]
        timestep = fstep = prevtime = fbest = 0.0
        istep = repeating = 0
        compute_new, pending = True, False
        while istep < nsteps and prevtime < T:
            if pending:
                push(cur, timestep, 1)
                pending = False
            tmin = prevtime
            if compute_new:
                if prevtime >= t_infl:
                    fbest = fstep - fdelta
                elif (fdelta > 0 and fstep + fdelta > f_infl) or (fdelta < 0 and fstep + fdelta < f_infl):
                    fbest = f_infl - (fstep + fdelta - f_infl)
                    tmin = t_infl - 1e-7
                else:
                    fbest = fstep + fdelta
            if abs(aL) <= eps_lin:
                if not abs(vL) > 1e-15:
                    return None
                timestep = fbest / vL
            else:
                fnew, found = fbest, False
                roots = _solvequad(aL * 0.5, vL, -fnew)
                if len(roots) == 0 and abs(fnew - largest_step) < 1e-7:
                    timestep, found = T, True
                if not found:
                    for r in roots:
                        if r > tmin and (not found or timestep > r):
                            timestep, found = r, True
                    if not found and tmin < t_infl:
                        fnew = fstep
                        for r in _solvequad(aL * 0.5, vL, -fnew):
                            if r > tmin and (not found or timestep > r):
                                timestep, fbest, found = r, fnew, True
                if not found:
                    return None
            prev = list(cur)
            dqscale = 1.0
            for i in range(n):
                dQ[i] = q0[i] + timestep * (dq0[i] + timestep * 0.5 * acc[i]) - cur[i]
                if abs(dQ[i]) > res[i] * 1.01:
                    tscaled, fs = 0.0, False
                    for r in _solvequad(acc[i] * 0.5, dq0[i], q0[i] - cur[i] - res[i]):
                        if tmin <= r <= timestep:
                            if not fs or r < tscaled:
                                tscaled = r
                            fs = True
                    sc = (tscaled - prevtime) / (timestep - prevtime) if (fs and timestep > prevtime) else abs(res[i] / dQ[i])
                    if sc < dqscale:
                        dqscale = sc
            if dqscale < 1:
                repeating += 1
                if dqscale <= 0.01 or repeating > nsteps * 2:
                    return None
                timestep = prevtime + (timestep - prevtime) * dqscale
                for i in range(n):
                    dQ[i] = q0[i] + timestep * (dq0[i] + timestep * 0.5 * acc[i]) - cur[i]
                    vel[i] = dq0[i] + timestep * acc[i]
            else:
                for i in range(n):
                    vel[i] = dq0[i] + timestep * acc[i]
                repeating = 0
                if timestep > T + 1e-7:
                    if istep + 1 >= nsteps:
                        break
                    return None
                elif timestep > T:
                    timestep = T
            for i in range(n):                                              # AddStatesWithLimitCheck
                cur[i] = min(max(cur[i] + dQ[i], lower[i]), upper[i])
            pending = True
            moved, npost = False, 1
            for i in range(n):
                f = abs(cur[i] - prev[i])
                if f > res[i] * 1.01:
                    npost = max(npost, int(f / res[i] + 0.9999))
                if f > 0.0001:
                    moved = True
            if npost > 1:
                inp = 1 / float(npost)
                step = [(cur[i] - prev[i]) * inp for i in range(n)]
                for _ in range(npost - 1):
                    for i in range(n):
                        prev[i] += step[i]
                    push(prev, float("nan"), 0)
            if (not moved) or (istep + 1 < nsteps and repeating > 2) or dqscale >= 1:
                fstep, compute_new = fbest, True
                istep += 1
            else:
                compute_new = False
            prevtime = timestep
        if abs(fstep - largest_step) > abs(fdelta):
            return None
        npost = 1
        for i in range(n):
            f = abs(q1[i] - cur[i])
            if f > res[i] * 1.01:
                npost = max(npost, int(f / res[i] + 0.9999))
        if npost > 1:
            inp = 1 / float(npost)
            step = [(q1[i] - cur[i]) * inp for i in range(n)]
            prev = list(cur)
            for _ in range(npost - 1):
                for i in range(n):
                    prev[i] += step[i]
                push(prev, float("nan"), 2)
        if pending and self._metric(cur, q1) > 1e-7:
            push(cur, timestep, 1)
            if check_end:
                filled[0] = 3                                               # q1 was checked but the walk stopped short of it (:3280)
        return pack()

    def _metric(self, a, b):
        """params->_distmetricfn (SimpleDistanceMetric, planningutils.cpp:4714-4723) with the robot's DOF weights."""
        d = self.diffstate(a, b)
        s = 0.0
        for i in range(len(d)):
            s += float(self.weights[i]) ** 2 * float(d[i]) * float(d[i])
        return math.sqrt(s)

    def CheckRamp(self, q0, q1, dq0, dq1, timeelapsed, interval=IT_Closed,
                  options=0xffff | CFO_CheckWithPerturbation | CFO_FillCheckedConfiguration, filterreturn=None):
        """Check(q0, q1, dq0, dq1, timeelapsed, ...) with valid velocities: every sample of the ramp in ONE CheckCollisionBatch."""
        from .checker import BATCH_ENV, BATCH_SELF
        opts = options & self.filtermask
        flags = (BATCH_ENV if opts & CFO_CheckEnvCollisions else 0) | (BATCH_SELF if opts & CFO_CheckSelfCollisions else 0)
        r = self.ramp_samples(q0, q1, dq0, dq1, timeelapsed, interval)
        if r is None:
            if filterreturn is not None:
                filterreturn.Clear()
                filterreturn._returncode = CFO_StateSettingError
            return CFO_StateSettingError
        s, t, filled = r
        mult = 3 if (opts & CFO_CheckWithPerturbation) and self.perturbation > 0 else 1
        codes = np.zeros(len(s), dtype=np.int64)
        if len(s):
            allq = self._perturbed(s) if mult == 3 else s
            v = self.checker.CheckCollisionBatch(self.robot, allq, flags=flags, active=self.active)["verdict"].reshape(len(s), mult)
            for i in range(len(s)):
                for c in range(mult):
                    code = self._code(int(v[i, c]))
                    if code:
                        codes[i] = code
                        break
        return self._finish(s, t, codes, options, filterreturn, filled=filled, end_first=(interval in (IT_OpenStart, IT_Closed)))

    def _perturbed(self, samples):
        """CFO_CheckWithPerturbation (:2344-2366): every state is also checked at +/- perturbation*resolution, clamped."""
        up = np.minimum(np.maximum(samples + self.perturbation * self.resolution, self.lower), self.upper)
        dn = np.minimum(np.maximum(samples - self.perturbation * self.resolution, self.lower), self.upper)
        return np.stack([samples, up, dn], axis=1).reshape(-1, samples.shape[1])

    @staticmethod
    def _code(v):
        return CFO_CheckEnvCollisions if (v & 1) else (CFO_CheckSelfCollisions if (v & 2) else 0)

    def _finish(self, samples, times, codes, options, filterreturn, filled=None, end_first=None):
        """First failing sample in reference order -> return code + ConstraintFilterReturn."""
        bad = np.nonzero(codes)[0]
        first = int(bad[0]) if len(bad) else -1
        if filterreturn is not None:
            filterreturn.Clear()
            if first >= 0:
                filterreturn._returncode = int(codes[first])
                filterreturn._invalidvalues = samples[first].tolist()
                filterreturn._fTimeWhenInvalid = float(times[first])
            if options & CFO_FillCheckedConfiguration:
                upto = len(samples) if first < 0 else first + 1
                order = list(range(upto))
                # time order: the end sample (checked first) is reported last when the whole edge is free (:3613-3624)
                lead = (times[0] == 1.0) if end_first is None else end_first
                if first < 0 and len(order) and lead and len(order) > 1:
                    order = order[1:] + [0]
                if filled is not None:
                    order = [i for i in order if filled[i] == 1]
                filterreturn._configurations = [samples[i].tolist() for i in order]
                filterreturn._configurationtimes = [float(times[i]) for i in order]
        return int(codes[first]) if first >= 0 else 0

    # ---- batched path (GPU): one call per edge, or one call for many edges
    def Check(self, q0, q1, interval=IT_Closed, options=0xffff | CFO_CheckWithPerturbation | CFO_FillCheckedConfiguration,
              filterreturn=None):
        return self.CheckMany([(q0, q1)], interval, options, [filterreturn])[0]

    def CheckMany(self, edges, interval=IT_Closed, options=0xffff | CFO_CheckWithPerturbation | CFO_FillCheckedConfiguration,
                  filterreturns=None):
        """All samples of all edges in ONE CheckCollisionBatch call (speculative: no early exit inside the batch)."""
        from .checker import BATCH_ENV, BATCH_SELF
        opts = options & self.filtermask
        flags = (BATCH_ENV if opts & CFO_CheckEnvCollisions else 0) | (BATCH_SELF if opts & CFO_CheckSelfCollisions else 0)
        per_edge, blocks = [], []
        mult = 3 if (opts & CFO_CheckWithPerturbation) and self.perturbation > 0 else 1
        for q0, q1 in edges:
            s, t, _ = self.edge_samples(q0, q1, interval)
            per_edge.append((s, t))
            blocks.append(self._perturbed(s) if mult == 3 and len(s) else s)
        allq = np.concatenate(blocks) if blocks else np.zeros((0, len(self.active)))
        if len(allq):
            verdict = self.checker.CheckCollisionBatch(self.robot, allq, flags=flags, active=self.active)["verdict"]
        else:
            verdict = np.zeros(0, dtype=np.uint32)
        rets, pos = [], 0
        for k, (s, t) in enumerate(per_edge):
            m = len(s)
            v = verdict[pos:pos + m * mult].reshape(m, mult) if m else np.zeros((0, mult), dtype=np.uint32)
            pos += m * mult
            codes = np.zeros(m, dtype=np.int64)
            for i in range(m):
                for c in range(mult):                                    # state, then +perturbation, then -perturbation
                    code = self._code(int(v[i, c]))
                    if code:
                        codes[i] = code
                        break
            fr = filterreturns[k] if filterreturns is not None else None
            rets.append(self._finish(s, t, codes, options, fr))
        return rets

    # ---- the literal per-sample loop (any CollisionCheckerBase-shaped checker; used as the comparison arm)
    def CheckSerial(self, q0, q1, interval=IT_Closed, options=0xffff | CFO_CheckWithPerturbation | CFO_FillCheckedConfiguration,
                    filterreturn=None, env=None):
        opts = options & self.filtermask
        env = env or self.robot.GetEnv()
        s, t, _ = self.edge_samples(q0, q1, interval)
        saved = self.robot.GetDOFValues()
        codes = np.zeros(len(s), dtype=np.int64)
        try:
            for i in range(len(s)):
                states = self._perturbed(s[i:i + 1]) if (opts & CFO_CheckWithPerturbation) and self.perturbation > 0 else s[i:i + 1]
                for q in states:
                    self.robot.SetDOFValues(q.tolist(), self.active)
                    if (opts & CFO_CheckEnvCollisions) and env.CheckCollision(self.robot):
                        codes[i] = CFO_CheckEnvCollisions
                        break
                    if (opts & CFO_CheckSelfCollisions) and self.robot.CheckSelfCollision():
                        codes[i] = CFO_CheckSelfCollisions
                        break
                if codes[i]:
                    break                                                  # the reference returns at the first failure
        finally:
            self.robot.SetDOFValues(saved)
        return self._finish(s, t, codes, options, filterreturn)


class BatchedPathConstraint(DynamicsCollisionConstraint):
    """Twin of plugins/cudacollision/cudabatched.h::BatchedPathConstraint — what PlannerParameters::_checkpathvelocityconstraintsfn
    is rebound to so that UNMODIFIED stock planners batch every edge: same call signature and ConstraintFilterReturn as
    DynamicsCollisionConstraint.Check, but the edge goes to the GPU in one cc_planner_check_edges call (sample generation,
    +/- perturbation copies, first failing sample, return code all inside the library)."""

    def __init__(self, robot, checker, perturbation=0.1):
        DynamicsCollisionConstraint.__init__(self, robot, checker, perturbation=perturbation)
        self._planner = BirrtPlanner(robot, checker)
        self.edges = 0
        self.samples = 0

    def Check(self, q0, q1, interval=IT_Closed, options=0xffff | CFO_CheckWithPerturbation | CFO_FillCheckedConfiguration,
              filterreturn=None, dq0=None, dq1=None, timeelapsed=0.0):
        P = self._planner
        ramp = timeelapsed > 0 and dq0 is not None and dq1 is not None
        opts = options & (CFO_CheckEnvCollisions | CFO_CheckSelfCollisions | CFO_CheckWithPerturbation)
        r = P.CheckEdges(q0, q1, interval, opts, self.perturbation, dq0 if ramp else None, dq1 if ramp else None,
                         [timeelapsed] if ramp else None)[0]
        self.edges += 1
        self.samples += max(r["nsamples"], 0)
        code = r["return_code"]
        if filterreturn is not None:
            filterreturn.Clear()
            filterreturn._returncode = code
            first = r["first_invalid"]
            if code and first >= 0:
                filterreturn._invalidvalues = r["invalid_values"].tolist()
                filterreturn._fTimeWhenInvalid = r["time_when_invalid"]
            if (options & CFO_FillCheckedConfiguration) and code != CFO_StateSettingError:
                if ramp:
                    s, t, kinds = P.RampSamples(q0, q1, dq0, dq1, timeelapsed, interval)
                    endfirst = interval in (IT_OpenStart, IT_Closed)
                else:
                    s, t = P.EdgeSamples(q0, q1, interval), P.EdgeSampleTimes(q0, q1, interval)
                    kinds = np.ones(len(s), dtype=np.int64)
                    endfirst = len(t) > 0 and t[0] == 1.0
                upto = len(s) if first < 0 else min(len(s), first + 1)
                order = list(range(upto))
                if first < 0 and endfirst and len(order) > 1:
                    order = order[1:] + [0]
                order = [i for i in order if kinds[i] == 1]
                filterreturn._configurations = [s[i].tolist() for i in order]
                filterreturn._configurationtimes = [float(t[i]) for i in order]
        return code


def CheckPathAllConstraints(constraint, q0, q1, interval=IT_Closed, options=0xffff, filterreturn=None):
    """PlannerParameters::CheckPathAllConstraints (planner.h:373-379) for positions-only edges."""
    return constraint.Check(q0, q1, interval, options, filterreturn)


# ---------------------------------------------------------------------------------------------------------------------
class ShortcutLinearPlanner(object):
    """RaveCreatePlanner(env, 'shortcut_linear') (plugins/rplanners/linearshortcutadvanced.cpp), literally: one
    CheckPathAllConstraints call per candidate segment, in the reference's order, over a DynamicsCollisionConstraint.  The
    library runs the same optimiser with whole shortcuts per checker call (csrc/cc_planner.cpp, CC_POST_SHORTCUT_LINEAR);
    this class is the per-call form the tests compare it with.  Retiming (the `_linearretimer` that follows) is not part of
    the collision path."""

    def __init__(self, constraint, weights=None, step_length=0.04, max_iterations=100, seed=0):
        self.c = constraint
        n = len(constraint.active)
        self.w2 = np.ones(n) if weights is None else np.asarray(weights, dtype=np.float64) ** 2
        self.step_length = step_length if step_length > 0 else 0.04          # _InitPlan (:52-59)
        self.max_iterations = max_iterations if max_iterations > 0 else 100
        bg = np.random.MT19937()
        bg._legacy_seeding(int(seed))                                        # mt19937ar.h init_genrand(seed)
        self._bg = bg
        self.calls = 0                                                       # CheckPathAllConstraints calls made

    def _u32(self):
        return int(self._bg.random_raw())

    def _dist(self, a, b):
        """SimpleDistanceMetric::Eval (planningutils.cpp:4714-4723), summed in DOF order."""
        d = self.c.diffstate(a, b)
        s = 0.0
        for i in range(len(d)):
            s += float(self.w2[i]) * float(d[i]) * float(d[i])
        return math.sqrt(s)

    def _check(self, q0, q1, interval):
        fr = ConstraintFilterReturn()
        self.calls += 1
        try:
            ret = self.c.Check(q0, q1, interval, 0xffff | CFO_FillCheckedConfiguration, fr)
        except NotImplementedError:
            return CFO_StateSettingError, []
        return ret, [np.array(c, dtype=np.float64) for c in fr._configurations]

    def _subsample(self, waypoints):
        """_SubsampleTrajectory (:578-652)."""
        c = self.c
        q0 = np.array(waypoints[0], dtype=np.float64)
        path = [[q0.copy(), 0.0]]
        qcur = q0.copy()
        for ip in range(1, len(waypoints)):
            q1 = np.array(waypoints[ip], dtype=np.float64)
            dq = c.diffstate(q1, q0)
            nsteps = 1
            for i in range(len(dq)):
                steps = int(abs(dq[i]) / c.resolution[i]) if c.resolution[i] != 0 else int(abs(dq[i]) * 100)
                nsteps = max(nsteps, steps)
            dq = dq * (1.0 / nsteps)
            for f in range(1, nsteps):
                qn, dev = c.neighstate(qcur, dq)
                if dev:                                                      # NSS_SuccessfulWithDeviation: stop subsampling
                    qcur = path[-1][0].copy()
                    break
                qcur = qn
                path.append([qcur.copy(), self._dist(path[-1][0], qcur)])
            path.append([q1.copy(), self._dist(path[-1][0], q1)])
            qcur = q1.copy()
            q0 = q1
        qlast = np.array(waypoints[-1], dtype=np.float64)
        path.append([qlast, self._dist(path[-1][0], qlast)])                 # the reference adds the last point twice (:648-650)
        return path

    def _optimize_path(self, path):
        """_OptimizePath (:188-339)."""
        num = 0
        iiter = self.max_iterations
        while iiter > 0 and len(path) > 2:
            iiter -= 1
            end = 2 + self._u32() % (len(path) - 2)
            start = self._u32() % (end - 1)
            total = 0.0
            for j in range(start + 1, end + 1):                              # plain left-to-right sum like the reference
                total += path[j][1]
            if self._dist(path[start][0], path[end][0]) > total - 0.1 * self.step_length:
                continue
            ret, cfgs = self._check(path[start][0], path[end][0], IT_Open)
            if ret != 0 or len(cfgs) == 0:
                continue
            dists = [self._dist(path[start][0], cfgs[0])]
            for k in range(1, len(cfgs)):
                dists.append(self._dist(cfgs[k - 1], cfgs[k]))
            dists.append(self._dist(cfgs[-1], path[end][0]))
            newtotal = 0.0
            for d in dists:
                newtotal += d
            if newtotal > total - 0.1 * self.step_length:
                continue
            path[end][1] = dists[-1]
            path[start + 1:end] = [[cfgs[k], dists[k]] for k in range(len(cfgs))]
            num += 1
        return num

    def _optimize_path_one_dof(self, path):
        """_OptimizePathOneDOF (:344-569)."""
        c = self.c
        n = len(c.active)
        num = 0
        for _ in range(self.max_iterations):
            if len(path) <= 2:
                break
            idof = min(self._u32() % (n + 1), n - 1)
            end = 2 + self._u32() % (len(path) - 2)
            start = self._u32() % (end - 1)
            if end == start + 1:
                continue
            totaldof, total, curv = 0.0, 0.0, path[start][0][idof]
            for j in range(start + 1, end + 1):
                totaldof += abs(path[j][0][idof] - curv)
                total += path[j][1]
                curv = path[j][0][idof]
            expected = path[end][0][idof] - path[start][0][idof]
            if abs(expected) > totaldof - 0.1 * c.resolution[idof]:
                continue
            sp, ok = [], True
            fdelta, fstart, fcur = expected / total, path[start][0][idof], 0.0
            newdof = newtotal = 0.0
            vcur = path[start][0].copy()
            for j in range(start + 1, end + 1):
                fcur += path[j][1]
                vnext = path[j][0].copy()
                if j != end:
                    vnext[idof] = fstart + fcur * fdelta
                base = sp[-1][0] if sp else vcur
                vnext, dev = c.neighstate(base.copy(), vnext - base)
                if j == end and dev:
                    ok = False
                    break
                ret, cfgs = self._check(vcur, vnext, IT_OpenStart)
                if ret != 0 or len(cfgs) == 0:
                    ok = False
                    break
                sp.append([cfgs[0], self._dist(path[start][0], cfgs[0])])   # sic: from the start node for every segment (:506-508)
                newdof += abs(cfgs[0][idof] - path[start][0][idof])
                newtotal += sp[-1][1]
                for k in range(1, len(cfgs)):
                    newdof += abs(cfgs[k][idof] - sp[-1][0][idof])
                    if newdof > totaldof - 0.1 * c.resolution[idof]:
                        ok = False
                        break
                    sp.append([cfgs[k], self._dist(sp[-1][0], cfgs[k])])
                    newtotal += sp[-1][1]
                    if newtotal > 1.1 * total:
                        ok = False
                        break
                if not ok:
                    break
                vcur = vnext
            if not ok or not sp:
                continue
            num += 1
            path[start + 1:end + 1] = sp
        return num

    def PlanPath(self, waypoints):
        """Returns (new waypoints [m, ndof], shortcuts of pass 1, shortcuts of pass 2)."""
        waypoints = np.asarray(waypoints, dtype=np.float64)
        if len(waypoints) < 2:
            return waypoints.copy(), 0, 0
        path = self._subsample(waypoints)
        a = self._optimize_path(path)
        b = self._optimize_path_one_dof(path)
        return np.array([p[0] for p in path]), a, b


# ---------------------------------------------------------------------------------------------------------------------
# BiRRT + linear shortcutting in lockstep over the batched entry (libcudacollision.so: csrc/cc_planner.cpp)

import ctypes


class cc_cspace(ctypes.Structure):
    _fields_ = [("ndof", ctypes.c_int32), ("reserved", ctypes.c_int32), ("lower", ctypes.c_void_p), ("upper", ctypes.c_void_p),
                ("resolution", ctypes.c_void_p), ("weights", ctypes.c_void_p), ("circular", ctypes.c_void_p)]


class cc_birrt_params(ctypes.Structure):
    _fields_ = [("step_length", ctypes.c_double), ("goal_bias", ctypes.c_double), ("point_tolerance", ctypes.c_double),
                ("max_iterations", ctypes.c_int32), ("path_optimize_iterations", ctypes.c_int32),
                ("smoother_iterations", ctypes.c_int32), ("max_extend_steps", ctypes.c_int32), ("speculate_steps", ctypes.c_int32),
                ("post_processor", ctypes.c_int32), ("minimum_goal_paths", ctypes.c_int32), ("max_planning_time_ms", ctypes.c_int32)]


class cc_plan_result(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("iterations", ctypes.c_int32), ("nodes_forward", ctypes.c_int32),
                ("nodes_backward", ctypes.c_int32), ("waypoints_raw", ctypes.c_int32), ("waypoints", ctypes.c_int32),
                ("rounds", ctypes.c_int32), ("shortcuts", ctypes.c_int32), ("length_raw", ctypes.c_double),
                ("length", ctypes.c_double), ("configs_submitted", ctypes.c_uint64), ("configs_serial", ctypes.c_uint64),
                ("start_index", ctypes.c_int32), ("goal_index", ctypes.c_int32), ("goal_paths", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class cc_edge_result(ctypes.Structure):
    _fields_ = [("return_code", ctypes.c_int32), ("first_invalid", ctypes.c_int32), ("nsamples", ctypes.c_int32),
                ("perturbed", ctypes.c_int32), ("time_when_invalid", ctypes.c_double)]


PLAN_FAILED, PLAN_SOLVED, PLAN_BAD_INITIAL, PLAN_BAD_GOAL, PLAN_INTERRUPTED = 0, 1, 2, 3, 4
PA_None, PA_Interrupt, PA_ReturnWithAnySolution = 0, 1, 2                # PlannerAction (planner.h:76-79)
POST_PROCESSORS = {"linearsmoother": 0, "shortcut_linear": 1, "linearsmoother 0": 2}          # PlannerParameters::_sPostProcessingPlanner
PLANNER_EXPORTS = ["cc_planner_create", "cc_planner_destroy", "cc_planner_set_query", "cc_planner_set_checker", "cc_planner_set_samplers", "cc_planner_set_progress_callback",
                   "cc_planner_edge_samples", "cc_planner_edge_sample_times", "cc_planner_ramp_samples", "cc_planner_check_edges", "cc_birrt_plan_batch", "cc_birrt_plan_batch_multi", "cc_basicrrt_plan_batch", "cc_planner_get_path", "cc_planner_get_totals"]


def _planner_lib():
    from .checker import load_library
    lib = load_library()
    if getattr(lib, "_planner_ready", False):
        return lib
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.cc_planner_create.argtypes = [ctypes.POINTER(cc_cspace), ctypes.POINTER(vp)]
    lib.cc_planner_destroy.argtypes = [vp]
    lib.cc_planner_set_query.argtypes = [vp, vp]
    lib.cc_planner_set_checker.argtypes = [vp, vp, vp]
    lib.cc_planner_edge_samples.argtypes = [vp, vp, vp, ci, vp, ci, ctypes.POINTER(ci)]
    lib.cc_planner_edge_sample_times.argtypes = [vp, vp, vp, ci, vp, ci, ctypes.POINTER(ci)]
    lib.cc_planner_ramp_samples.argtypes = [vp, vp, vp, vp, vp, ctypes.c_double, ci, vp, vp, vp, ci, ctypes.POINTER(ci)]
    lib.cc_planner_check_edges.argtypes = [vp, ci, vp, vp, vp, vp, vp, ci, ci, ctypes.c_double, vp, vp]
    lib.cc_birrt_plan_batch.argtypes = [vp, ctypes.POINTER(cc_birrt_params), ci, vp, vp, vp, vp]
    lib.cc_basicrrt_plan_batch.argtypes = [vp, ctypes.POINTER(cc_birrt_params), ci, ci, vp, vp, vp, vp]
    lib.cc_birrt_plan_batch_multi.argtypes = [vp, ctypes.POINTER(cc_birrt_params), ci, vp, vp, vp, vp, vp, vp]
    lib.cc_planner_get_path.argtypes = [vp, ci, ci, vp, ci, ctypes.POINTER(ci)]
    lib.cc_planner_get_totals.argtypes = [vp, vp]
    lib.cc_planner_set_samplers.argtypes = [vp, vp, vp, vp]
    lib.cc_planner_set_progress_callback.argtypes = [vp, vp, vp]
    lib._planner_ready = True
    return lib


class BirrtPlanner(object):
    """RaveCreatePlanner(env, 'birrt') on the batched collision entry, for one robot over its active DOFs.

    `checker` is a CudaCollisionChecker (edge checks go through cc_check_batch) or a pair (function pointer, user block)
    with the cc_group_check_fn signature (the tests plug the oracle's serial loop in that way).  `PlanBatch` runs many
    (start, goal) problems in lockstep; `PlanPath` is the single-problem call of the reference API."""

    def __init__(self, robot, checker, step_length=0.04, max_iterations=10000, goal_bias=0.01, smoother_iterations=100,
                 path_optimize_iterations=10, point_tolerance=0.2, speculate_steps=-8, post_processor="linearsmoother",
                 minimum_goal_paths=1, max_planning_time_ms=0):
        self.lib = _planner_lib()
        self.robot = robot
        self.active = list(robot.GetActiveDOFIndices())
        joints = [robot.joints[d] for d in self.active]
        self._lo = np.array([j.lower for j in joints], dtype=np.float64)
        self._hi = np.array([j.upper for j in joints], dtype=np.float64)
        self._res = np.array([j.resolution for j in joints], dtype=np.float64)
        self._w = np.array([j.weight for j in joints], dtype=np.float64)
        self._circ = np.array([1 if j.circular else 0 for j in joints], dtype=np.uint8)
        self.ndof = len(self.active)
        sp = cc_cspace(self.ndof, 0, self._lo.ctypes.data, self._hi.ctypes.data, self._res.ctypes.data, self._w.ctypes.data,
                       self._circ.ctypes.data)
        h = ctypes.c_void_p()
        self._ck(self.lib.cc_planner_create(ctypes.byref(sp), ctypes.byref(h)))
        self.handle = h
        self.params = cc_birrt_params(step_length, goal_bias, point_tolerance, max_iterations, path_optimize_iterations,
                                      smoother_iterations, 100, speculate_steps, POST_PROCESSORS[post_processor],
                                      int(minimum_goal_paths), int(max_planning_time_ms))     # RRTParameters::_minimumgoalpaths, _nMaxPlanningTime
        self._keep = None
        if isinstance(checker, tuple):
            fn, blk = checker
            self._keep = (fn, blk)
            self._ck(self.lib.cc_planner_set_checker(h, fn, ctypes.addressof(blk)))
        else:
            bq = checker.bind(robot, self.active)
            self._keep = bq
            self._ck(self.lib.cc_planner_set_query(h, bq.handle))

    def _ck(self, rc):
        if rc != 0:
            raise RuntimeError("cudacollision planner error %d: %s" % (rc, (self.lib.cc_last_error() or b"").decode()))

    def __del__(self):
        try:
            if self.handle:
                self.lib.cc_planner_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def EdgeSamples(self, q0, q1, interval=IT_Closed):
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        q1 = np.ascontiguousarray(q1, dtype=np.float64)
        cnt = ctypes.c_int(0)
        self._ck(self.lib.cc_planner_edge_samples(self.handle, q0.ctypes.data, q1.ctypes.data, interval, None, 0, ctypes.byref(cnt)))
        if cnt.value < 0:
            return None
        out = np.zeros((max(cnt.value, 1), self.ndof))
        self._ck(self.lib.cc_planner_edge_samples(self.handle, q0.ctypes.data, q1.ctypes.data, interval, out.ctypes.data,
                                                  cnt.value, ctypes.byref(cnt)))
        return out[:cnt.value]

    def EdgeSampleTimes(self, q0, q1, interval=IT_Closed):
        """ConstraintFilterReturn::_configurationtimes of EdgeSamples' rows (cc_planner_edge_sample_times)."""
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        q1 = np.ascontiguousarray(q1, dtype=np.float64)
        cnt = ctypes.c_int(0)
        self._ck(self.lib.cc_planner_edge_sample_times(self.handle, q0.ctypes.data, q1.ctypes.data, interval, None, 0, ctypes.byref(cnt)))
        if cnt.value < 0:
            return None
        out = np.zeros(max(cnt.value, 1))
        self._ck(self.lib.cc_planner_edge_sample_times(self.handle, q0.ctypes.data, q1.ctypes.data, interval, out.ctypes.data,
                                                       cnt.value, ctypes.byref(cnt)))
        return out[:cnt.value]

    def RampSamples(self, q0, q1, dq0, dq1, timeelapsed, interval=IT_Closed):
        """(samples, times, filled) of cc_planner_ramp_samples, or None for CFO_StateSettingError."""
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (q0, q1, dq0, dq1)]
        cnt = ctypes.c_int(0)
        args = [self.handle] + [x.ctypes.data for x in a] + [float(timeelapsed), interval]
        self._ck(self.lib.cc_planner_ramp_samples(*(args + [None, None, None, 0, ctypes.byref(cnt)])))
        if cnt.value < 0:
            return None
        m = max(cnt.value, 1)
        out, t, f = np.zeros((m, self.ndof)), np.zeros(m), np.zeros(m, dtype=np.uint8)
        self._ck(self.lib.cc_planner_ramp_samples(*(args + [out.ctypes.data, t.ctypes.data, f.ctypes.data, cnt.value, ctypes.byref(cnt)])))
        return out[:cnt.value], t[:cnt.value], f[:cnt.value].astype(np.int64)

    def CheckEdges(self, q0, q1, interval=IT_Closed, options=0xffff, perturbation=0.1, dq0=None, dq1=None, timeelapsed=None):
        """cc_planner_check_edges: CheckPathAllConstraints for many edges in one checker call.  Returns a list of dicts
        (return_code, first_invalid, nsamples, perturbed, time_when_invalid, invalid_values)."""
        q0 = np.ascontiguousarray(q0, dtype=np.float64).reshape(-1, self.ndof)
        q1 = np.ascontiguousarray(q1, dtype=np.float64).reshape(-1, self.ndof)
        k = q0.shape[0]
        ptr = lambda a: None if a is None else a.ctypes.data
        if dq0 is not None:
            dq0 = np.ascontiguousarray(dq0, dtype=np.float64).reshape(-1, self.ndof)
            dq1 = np.ascontiguousarray(dq1, dtype=np.float64).reshape(-1, self.ndof)
            timeelapsed = np.ascontiguousarray(timeelapsed, dtype=np.float64).reshape(-1)
        res = (cc_edge_result * max(k, 1))()
        inv = np.zeros((max(k, 1), self.ndof))
        self._ck(self.lib.cc_planner_check_edges(self.handle, k, q0.ctypes.data, q1.ctypes.data, ptr(dq0), ptr(dq1), ptr(timeelapsed),
                                                 interval, options & 0xffffffff, float(perturbation), ctypes.addressof(res), inv.ctypes.data))
        out = []
        for i in range(k):
            d = dict((f, getattr(res[i], f)) for f, _ in cc_edge_result._fields_)
            d["return_code"] &= 0xffffffff
            d["invalid_values"] = inv[i].copy() if d["first_invalid"] >= 0 else None
            out.append(d)
        return out

    def PlanBatch(self, starts, goals, seeds=None, planner="birrt", extent_type=0):
        """Returns a list of dicts (status, iterations, path [w, ndof], path_raw, lengths, config counts) + totals.
        planner: "birrt" (BirrtPlanner) or "basicrrt" (BasicRrtPlanner, extent_type = _nRRTExtentType)."""
        starts = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, self.ndof)
        goals = np.ascontiguousarray(goals, dtype=np.float64).reshape(-1, self.ndof)
        k = starts.shape[0]
        seeds = np.zeros(k, dtype=np.uint32) if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        res = (cc_plan_result * max(k, 1))()
        if planner == "basicrrt":
            self._ck(self.lib.cc_basicrrt_plan_batch(self.handle, ctypes.byref(self.params), extent_type, k, starts.ctypes.data,
                                                     goals.ctypes.data, seeds.ctypes.data, ctypes.addressof(res)))
        else:
            self._ck(self.lib.cc_birrt_plan_batch(self.handle, ctypes.byref(self.params), k, starts.ctypes.data, goals.ctypes.data,
                                                  seeds.ctypes.data, ctypes.addressof(res)))
        return self._collect(res, k)

    def _collect(self, res, k):
        out = []
        for i in range(k):
            r = res[i]
            d = dict((f, getattr(r, f)) for f, _ in cc_plan_result._fields_ if f != "reserved")
            for which, key in ((0, "path"), (1, "path_raw")):
                w = ctypes.c_int(0)
                self._ck(self.lib.cc_planner_get_path(self.handle, i, which, None, 0, ctypes.byref(w)))
                p = np.zeros((max(w.value, 1), self.ndof))
                self._ck(self.lib.cc_planner_get_path(self.handle, i, which, p.ctypes.data, w.value, ctypes.byref(w)))
                d[key] = p[:w.value]
            out.append(d)
        tot = np.zeros(4, dtype=np.uint64)
        self._ck(self.lib.cc_planner_get_totals(self.handle, tot.ctypes.data))
        self.totals = dict(rounds=int(tot[0]), checker_calls=int(tot[1]), configs_submitted=int(tot[2]), configs_serial=int(tot[3]))
        return out

    def PlanPath(self, start, goal, seed=0):
        return self.PlanBatch([start], [goal], [seed])[0]

    def SetProgressCallback(self, fn=None):
        """PlannerBase::RegisterPlanCallback: fn(problem, iteration) -> PA_None / PA_Interrupt / PA_ReturnWithAnySolution, asked at
        the top of every BiRRT iteration (cc_planner_set_progress_callback)."""
        proto = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int)
        self._progress = proto(lambda user, problem, iteration: int(fn(problem, iteration) or 0)) if fn is not None else None
        self._ck(self.lib.cc_planner_set_progress_callback(self.handle, ctypes.cast(self._progress, ctypes.c_void_p) if self._progress else None, None))

    def SetSamplers(self, samplegoalfn=None, sampleinitialfn=None):
        """PlannerParameters::_samplegoalfn / _sampleinitialfn: callables (problem index) -> configuration or None, asked once per
        main-loop iteration of every BiRRT problem (cc_planner_set_samplers)."""
        proto = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double))

        def wrap(fn):
            if fn is None:
                return None

            def cb(user, problem, ndof, out):
                v = fn(problem)
                if v is None:
                    return 0
                for k in range(ndof):
                    out[k] = float(v[k])
                return 1
            return proto(cb)
        self._samplers = (wrap(samplegoalfn), wrap(sampleinitialfn))          # keep the thunks alive
        g, i = self._samplers
        self._ck(self.lib.cc_planner_set_samplers(self.handle, ctypes.cast(g, ctypes.c_void_p) if g else None,
                                                  ctypes.cast(i, ctypes.c_void_p) if i else None, None))

    def PlanBatchMulti(self, starts_per_problem, goals_per_problem, seeds=None):
        """BiRRT with several initial / goal configurations per problem (lists of [k_i, ndof] arrays): cc_birrt_plan_batch_multi.
        The result dicts carry start_index / goal_index (GOALPATH::startindex / goalindex)."""
        k = len(starts_per_problem)
        S = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1, self.ndof) for a in starts_per_problem]
        G = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1, self.ndof) for a in goals_per_problem]
        sb = np.concatenate([[0], np.cumsum([len(a) for a in S])]).astype(np.int32)
        gb = np.concatenate([[0], np.cumsum([len(a) for a in G])]).astype(np.int32)
        sa = np.ascontiguousarray(np.concatenate(S)) if k else np.zeros((0, self.ndof))
        ga = np.ascontiguousarray(np.concatenate(G)) if k else np.zeros((0, self.ndof))
        seeds = np.zeros(k, dtype=np.uint32) if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint32)
        res = (cc_plan_result * max(k, 1))()
        self._ck(self.lib.cc_birrt_plan_batch_multi(self.handle, ctypes.byref(self.params), k, sb.ctypes.data, sa.ctypes.data,
                                                    gb.ctypes.data, ga.ctypes.data, seeds.ctypes.data, ctypes.addressof(res)))
        return self._collect(res, k)
