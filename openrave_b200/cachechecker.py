"""`CacheChecker` on the batched entry — host-side mirror of plugins/configurationcache/cachechecker.cpp.

The reference wraps any collision checker: `CheckCollision(robot)` / `CheckStandaloneSelfCollision(robot)` first ask a
configuration cache (a cover tree over weighted DOF vectors, configurationcachetree.cpp) whether a known colliding
configuration lies within `_collisionthresh` (answer: collision, with that node's link pair) or a known free one within
`_freespacethresh` (answer: free); only a miss reaches the wrapped checker, and its result is inserted unless a node already
sits within `thresh * _insertiondistancemult` (cachechecker.cpp:193-260, 309-360; configurationcachetree.cpp:364-440,
1203-1205, 1262-1274, 1324-1345).

This mirror keeps those rules and adds the batched form: every configuration of a batch is looked up against the cache as it
stood when the call came in, all misses go to the wrapped checker in ONE `CheckCollisionBatch`, and the results are inserted
in order under the same separation rule.  Differences from the serial wrapper, by design: a configuration never hits on a node
inserted earlier in the same batch (so the batch returns exact answers where the serial loop would have returned cached
ones), and the nearest-neighbour structure is an exact scan (same verdict as the cover tree, which returns the first collision
node it meets inside the threshold; the reported link pair may be another node's).  The cache answers are approximate by
construction — the thresholds are the caller's accuracy knob, exactly as in the reference.
"""
import numpy as np

from .checker import BATCH_ENV, BATCH_SELF


class ConfigurationCache(object):
    """ConfigurationCache + CacheTree semantics (configurationcachetree.cpp:1132-1345) with an exact scan."""

    def __init__(self, robot, active=None, collisionthresh=1.0, freespacethresh=0.2, insertiondistancemult=0.5):
        self.active = list(robot.GetActiveDOFIndices() if active is None else active)
        res = np.array([robot.joints[d].resolution for d in self.active], dtype=np.float64)
        self.weights = np.where(res > 0, 1.0 / np.where(res > 0, res, 1.0), 100.0)        # :1194-1201
        self.collisionthresh, self.freespacethresh = collisionthresh, freespacethresh     # :1203-1204
        self.insertiondistancemult = insertiondistancemult                                # :1205
        self.Reset()

    def Reset(self):
        n = len(self.active)
        self.conf = np.zeros((0, n))
        self.incollision = np.zeros(0, dtype=bool)
        self.pairs = []                                  # per node: (robot link, colliding link) or None

    def GetNumKnownNodes(self):
        return len(self.conf)

    def _dist(self, q):
        """[len(q), nodes] weighted distances (CacheTree::_ComputeDistance2, :251-259)."""
        d = (q[:, None, :] - self.conf[None, :, :]) * self.weights[None, None, :]
        return np.sqrt((d * d).sum(-1))

    def CheckCollisionMany(self, q):
        """ConfigurationCache::CheckCollision (:1324-1345) for every row: returns (ret, node) with ret 1 collision / 0 free / -1 miss."""
        q = np.asarray(q, dtype=np.float64).reshape(-1, len(self.active))
        ret = np.full(len(q), -1, dtype=np.int64)
        node = np.full(len(q), -1, dtype=np.int64)
        if len(self.conf) == 0 or len(q) == 0:
            return ret, node
        d = self._dist(q)
        dc = np.where(self.incollision[None, :], d, np.inf)
        df = np.where(~self.incollision[None, :], d, np.inf)
        ic, jf = dc.argmin(1), df.argmin(1)
        hitc = dc[np.arange(len(q)), ic] <= self.collisionthresh
        hitf = df[np.arange(len(q)), jf] <= self.freespacethresh
        ret[hitf], node[hitf] = 0, jf[hitf]
        ret[hitc], node[hitc] = 1, ic[hitc]              # a collision node inside its threshold wins (:384-388, 409-413)
        return ret, node

    def InsertConfiguration(self, conf, pair):
        """InsertConfiguration (:1262-1274): pair = (robot link, other link) for a collision, None for a free configuration;
        refused (False) when a node already lies within thresh * insertiondistancemult (CacheTree::_Insert, :503-505)."""
        conf = np.asarray(conf, dtype=np.float64).reshape(1, -1)
        minsep = (self.collisionthresh if pair is not None else self.freespacethresh) * self.insertiondistancemult
        if len(self.conf) and self._dist(conf).min() < minsep:
            return False
        self.conf = np.vstack([self.conf, conf])
        self.incollision = np.append(self.incollision, pair is not None)
        self.pairs.append(pair)
        return True


class CacheCollisionChecker(object):
    """RaveCreateCollisionChecker(env, 'CacheChecker <inner>') for one robot: environment and self caches in front of `inner`
    (any checker with the CollisionCheckerBase surface plus CheckCollisionBatch: the CUDA checker, or the oracle in tests)."""

    def __init__(self, inner, robot, **cacheargs):
        self.inner = inner
        self.robot = robot
        self.active = list(robot.GetActiveDOFIndices())
        self.cache = ConfigurationCache(robot, self.active, **cacheargs)          # environment collisions (cachechecker.cpp:193-260)
        self.selfcache = ConfigurationCache(robot, self.active, **cacheargs)      # self collisions (:309-360)
        self.stats = dict(checks=0, collisionhits=0, freehits=0, selfchecks=0, selfcollisionhits=0, selffreehits=0, inner_configs=0)

    # ---- everything the cache does not cover goes straight through (cachechecker.cpp:262-307)
    def __getattr__(self, name):
        return getattr(self.inner, name)

    def _q(self):
        v = self.robot.GetDOFValues()
        return np.array([v[d] for d in self.active], dtype=np.float64)

    def CheckCollision(self, a, b=None, report=None):
        if b is not None or a is not self.robot:
            return self.inner.CheckCollision(a, b, report=report) if b is not None else self.inner.CheckCollision(a, report=report)
        if report is not None:
            report.Reset()
        q = self._q()
        ret, node = self.cache.CheckCollisionMany(q[None])
        self.stats["checks"] += 1
        if ret[0] == 1:
            self.stats["collisionhits"] += 1
            if report is not None and self.cache.pairs[node[0]] is not None:
                report.AddLinkCollision(*self.cache.pairs[node[0]])
            return True
        if ret[0] == 0:
            self.stats["freehits"] += 1
            return False
        from .environment import CollisionReport
        rep = report if report is not None else CollisionReport()
        col = self.inner.CheckCollision(a, report=rep)
        self.stats["inner_configs"] += 1
        self.cache.InsertConfiguration(q, (rep.plink1, rep.plink2) if col else None)
        return col

    def CheckStandaloneSelfCollision(self, body, report=None):
        if body is not self.robot:
            return self.inner.CheckStandaloneSelfCollision(body, report=report)
        if report is not None:
            report.Reset()
        q = self._q()
        ret, node = self.selfcache.CheckCollisionMany(q[None])
        self.stats["selfchecks"] += 1
        if ret[0] == 1:
            self.stats["selfcollisionhits"] += 1
            if report is not None and self.selfcache.pairs[node[0]] is not None:
                report.AddLinkCollision(*self.selfcache.pairs[node[0]])
            return True
        if ret[0] == 0:
            self.stats["selffreehits"] += 1
            return False
        from .environment import CollisionReport
        rep = report if report is not None else CollisionReport()
        col = self.inner.CheckStandaloneSelfCollision(body, report=rep)
        self.stats["inner_configs"] += 1
        self.selfcache.InsertConfiguration(q, (rep.plink1, rep.plink2) if col else None)
        return col

    def CheckCollisionBatch(self, robot, q, flags=BATCH_ENV | BATCH_SELF, active=None):
        """The batched form: cache lookups for every row, ONE inner batch for the rows either cache misses, ordered inserts.
        Returns dict(verdict [n] (bit0 env, bit1 self), from_cache [n] (bit0 env answered by the cache, bit1 self))."""
        assert robot is self.robot and (active is None or list(active) == self.active)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, len(self.active))
        n = len(q)
        verdict = np.zeros(n, dtype=np.uint32)
        from_cache = np.zeros(n, dtype=np.uint32)
        need = np.zeros(n, dtype=bool)
        lookups = []
        for bit, fl, cache, keys in ((1, BATCH_ENV, self.cache, ("checks", "collisionhits", "freehits")),
                                     (2, BATCH_SELF, self.selfcache, ("selfchecks", "selfcollisionhits", "selffreehits"))):
            if not flags & fl:
                continue
            ret, _ = cache.CheckCollisionMany(q)
            self.stats[keys[0]] += n
            self.stats[keys[1]] += int((ret == 1).sum())
            self.stats[keys[2]] += int((ret == 0).sum())
            verdict[ret == 1] |= bit
            from_cache[ret >= 0] |= bit
            need |= ret < 0
            lookups.append((bit, cache, ret))
        idx = np.nonzero(need)[0]
        if len(idx):
            got = self.inner.CheckCollisionBatch(robot, q[idx], flags=flags, active=self.active)
            self.stats["inner_configs"] += len(idx)
            v = got["verdict"]
            for bit, cache, ret in lookups:
                first = got["first_env"] if bit == 1 else got["first_self"]
                for k, i in enumerate(idx):
                    if ret[i] >= 0:
                        continue                                             # this cache had an answer; the other one missed
                    hit = bool(v[k] & bit)
                    if hit:
                        verdict[i] |= bit
                    cache.InsertConfiguration(q[i], self._decode_pair(bit, first[k]) if hit else None)
        return dict(verdict=verdict, from_cache=from_cache)

    def _decode_pair(self, bit, rec):
        """first_env {link, other body id, other link} / first_self {link i, link j} of the inner batch -> (link1, link2) like the
        report of a single-state call; indices stay as they are when the inner checker cannot name the body."""
        links = self.robot.links
        try:
            if bit == 2:
                return (links[int(rec[0])], links[int(rec[1])])
            other = self.inner.body_from_id(int(rec[1]))
            return (links[int(rec[0])] if int(rec[0]) < len(links) else None, other.links[int(rec[2])])
        except Exception:
            return tuple(int(x) for x in rec)
