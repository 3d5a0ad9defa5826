"""Python binding of the CPU oracle (oracle/oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this
package.  The product (openrave_b200/) never does.

`OracleChecker(kind="port")` loads oracle/liboracle.so (the restated arithmetic);
`kind="reference"` loads oracle/_ref/liboracle_ref.so (same driver on the unmodified vendored PQP).
The class exposes the same checker surface as openrave_b200.checker.CudaCollisionChecker so the parity
tests run one body of code against both.
"""
import ctypes
import os
import subprocess

import numpy as np

from openrave_b200 import flatten
from openrave_b200.flatten import cc_joint

_HERE = os.path.dirname(os.path.abspath(__file__))

CO_ActiveDOFs = 0x10
BATCH_ENV, BATCH_SELF = 1, 2


class OrcOptions(ctypes.Structure):
    _fields_ = [("options", ctypes.c_int32), ("flags", ctypes.c_int32), ("report", ctypes.c_int32),
                ("want_band", ctypes.c_int32), ("band", ctypes.c_double), ("pose_mode", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


class OrcGroupChecker(ctypes.Structure):
    """`user` block of orc_group_check (oracle.cpp): the serial per-sample edge loop behind cc_group_check_fn."""
    _fields_ = [("scene", ctypes.c_void_p), ("active", ctypes.c_void_p), ("body", ctypes.c_int32), ("nactive", ctypes.c_int32),
                ("options", ctypes.c_int32), ("reserved", ctypes.c_int32), ("configs_checked", ctypes.c_uint64),
                ("seconds", ctypes.c_double)]


def build(ref=True):
    """Compile the oracle with its Makefile (idempotent).  The reference-linked variant needs
    /root/reference and is skipped silently where that tree does not exist (the GPU box uses the
    prebuilt oracle/_ref/liboracle_ref.so that travelled with the snapshot)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    if ref and os.path.isdir("/root/reference/plugins/pqprave/pqp"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def lib_path(kind):
    return os.path.join(_HERE, "liboracle.so") if kind == "port" else os.path.join(_HERE, "_ref", "liboracle_ref.so")


def available(kind):
    return os.path.isfile(lib_path(kind))


def _load(kind):
    path = lib_path(kind)
    if not os.path.isfile(path):
        if kind == "port":
            build(ref=False)
        else:
            raise OSError("%s missing: run `make -C oracle ref` where /root/reference exists" % path)
    lib = ctypes.CDLL(path)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    lib.orc_scene_create.restype = vp
    lib.orc_scene_destroy.argtypes = [vp]
    lib.orc_scene_add_body.argtypes = [vp, ctypes.c_char_p, ci, ci]
    lib.orc_body_set_link_mesh.argtypes = [vp, ci, ci, vp, ci, vp, ci]
    lib.orc_body_set_link_enabled.argtypes = [vp, ci, vp]
    lib.orc_body_set_link_transforms.argtypes = [vp, ci, vp]
    lib.orc_body_set_attached.argtypes = [vp, ci, ci, ci]
    lib.orc_body_set_grabbed.argtypes = [vp, ci, ci, ci, vp, vp, ci]
    lib.orc_body_release_grabbed.argtypes = [vp, ci, ci]
    lib.orc_body_set_removed.argtypes = [vp, ci, ci]
    lib.orc_body_set_kinematics.argtypes = [vp, ci, ci, ci, vp, vp, vp]
    lib.orc_body_set_nonadjacent_pairs.argtypes = [vp, ci, vp, ci]
    lib.orc_body_set_affine.argtypes = [vp, ci, ci, vp, vp]
    lib.orc_check_link_pairs.argtypes = [vp, ci, vp, vp, ci, vp]
    lib.orc_check_batch.argtypes = [vp, ci, vp, ci, ctypes.POINTER(OrcOptions), vp, ci, vp, vp, vp, vp, vp, vp, vp,
                                    ctypes.c_int64, vp, vp]
    lib.orc_tri_tri_collision.argtypes = [vp, vp, vp, vp]
    lib.orc_pair_queries.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci, cd, cd, vp, vp]
    lib.orc_pair_band.argtypes = [vp, ci, vp, ci, vp, ci, vp, ci, cd, vp]
    lib.orc_distance_batch.argtypes = [vp, ci, vp, ci, vp, vp, ci, cd, vp, vp]
    lib.orc_time_batch.restype = cd
    lib.orc_time_batch.argtypes = [vp, ci, vp, ci, ci, ci, vp, ci, vp]
    lib.orc_collide_pair.restype = ctypes.c_int64
    lib.orc_collide_pair.argtypes = [vp, ci, ci, vp, ci, ci, vp, ci, vp, vp, ctypes.c_int64]
    lib.orc_link_valid.argtypes = [vp, ci, ci]
    if kind != "port":
        lib.orc_ref_distance.restype = cd
        lib.orc_ref_distance.argtypes = [vp, ci, ci, vp, ci, ci, vp, cd, cd]
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class OracleChecker(object):
    """pqprave semantics on the CPU (collisionPQP.h), driven through the harness Environment."""

    def __init__(self, env=None, kind="port"):
        self.kind = kind
        self.lib = _load(kind)
        self.env = env
        self.scene = None
        self._ids = {}
        self._dirty = set()
        self._pose_only = set()      # bodies whose FK the C restatement cannot run (non-affine mimic): pose mode only
        self._options = 0

    # ---- CollisionCheckerBase surface
    def SetCollisionOptions(self, options):
        self._options = int(options)
        return True

    def GetCollisionOptions(self):
        return self._options

    def InitEnvironment(self):
        self.DestroyEnvironment()
        self.scene = self.lib.orc_scene_create()
        for b in self.env.GetBodies():
            self.InitKinBody(b)
        return True

    def DestroyEnvironment(self):
        if self.scene:
            self.lib.orc_scene_destroy(self.scene)
        self.scene = None
        self._ids = {}
        self._dirty = set()

    def __del__(self):
        try:
            self.DestroyEnvironment()
        except Exception:
            pass

    def InitKinBody(self, body):
        if id(body) in self._ids:
            return True
        bid = self.lib.orc_scene_add_body(self.scene, body.name.encode(), len(body.links), 1 if body.is_robot else 0)
        self._ids[id(body)] = bid
        for l in body.links:
            v, i = flatten.link_mesh_arrays(l)
            self.lib.orc_body_set_link_mesh(self.scene, bid, l.index, _ptr(v), v.shape[0], _ptr(i), i.shape[0])
        self._dirty.add(id(body))
        self._sync_body(body)
        return True

    def RemoveKinBody(self, body):
        bid = self._ids.pop(id(body), None)
        if bid is not None:
            self.lib.orc_body_set_removed(self.scene, bid, 1)

    def _on_body_changed(self, body, kind="structure"):
        self._dirty.add(id(body))

    def _sync_body(self, body):
        bid = self._ids[id(body)]
        T = flatten.transforms8(body.GetLinkTransformations())
        self.lib.orc_body_set_link_transforms(self.scene, bid, _ptr(T))
        en = np.array([1 if l.enabled else 0 for l in body.links], dtype=np.uint8)
        self.lib.orc_body_set_link_enabled(self.scene, bid, _ptr(en))

    def _sync(self):
        if not self.scene:
            self.InitEnvironment()
        for b in self.env.GetBodies():
            if id(b) not in self._ids:
                self.InitKinBody(b)
            if id(b) in self._dirty:
                self._sync_body(b)
        self._dirty.clear()
        epochs = self.__dict__.setdefault("_grab_epochs", {})
        for b in self.env.GetBodies():
            ep = getattr(b, "grab_epoch", 0)
            if epochs.get(id(b), 0) != ep:
                bid = self._ids[id(b)]
                self.lib.orc_body_release_grabbed(self.scene, bid, -1)
                for gid, link, rel, nc in flatten.grab_tables(b, lambda k: self._ids[id(k)]):
                    self.lib.orc_body_set_grabbed(self.scene, bid, gid, link, _ptr(rel), _ptr(nc) if len(nc) else None, len(nc))
                epochs[id(b)] = ep
                for g in b.grabbed:
                    if len(g.body.links) > 1:
                        self._sync_kinematics(g.body)

    def _sync_kinematics(self, body):
        bid = self._ids[id(body)]
        if body.GetDOF() > 0:
            try:
                arr, n = flatten.joint_table(body)
                self._pose_only.discard(bid)
            except NotImplementedError:
                # a mimic equation the C restatement's FK does not evaluate: such a body is checked in pose mode only (the
                # link poses then come from the harness' SetDOFValues, kinbody.fk_transforms)
                self._pose_only.add(bid)
                arr = None
            if arr is not None:
                aff = flatten.affects_table(body)
                cur = np.array(body.GetDOFValues(), dtype=np.float64)
                self.lib.orc_body_set_kinematics(self.scene, bid, body.GetDOF(), n, ctypes.addressof(arr), _ptr(aff), _ptr(cur))
        if len(body.links) > 1:
            na = np.array(body.GetNonAdjacentLinks(0, checker=self), dtype=np.int32)
        else:
            na = np.zeros(0, dtype=np.int32)
        self.lib.orc_body_set_nonadjacent_pairs(self.scene, bid, _ptr(na), int(na.size))
        mask = int(getattr(body, "affine_dofs", 0))
        axis = np.array(getattr(body, "affine_axis", (0.0, 0.0, 1.0)), dtype=np.float64)
        qstart = np.array(getattr(body, "affine_quat_start", (1.0, 0.0, 0.0, 0.0)), dtype=np.float64)
        self.lib.orc_body_set_affine(self.scene, bid, mask, _ptr(axis), _ptr(qstart))
        return bid

    def _check_link_pairs_at(self, body, transforms, pairs):
        """CheckCollision(link,link) for `pairs` with the body's links placed at `transforms`."""
        self._sync()
        bid = self._ids[id(body)]
        T = flatten.transforms8(transforms)
        packed = np.array([a | (b << 16) for a, b in pairs], dtype=np.int32)
        hit = np.zeros(len(pairs), dtype=np.uint8)
        if len(pairs):
            self.lib.orc_check_link_pairs(self.scene, bid, _ptr(T), _ptr(packed), len(pairs), _ptr(hit))
        return [bool(h) for h in hit]

    # ---- batched entry
    def check_batch(self, body, q, active=None, flags=BATCH_ENV | BATCH_SELF, report=False, want_band=False,
                    band=1e-4, want_fk=False, poses=None):
        """Returns dict(verdict, first_env[n,3], first_self[n,2], counters[n,2], penetration[n], near[n],
        env_pairs (report mode: [m,4] rows {config, link, other body, other link}), fk[n,nl,8])."""
        self._sync()
        bid = self._sync_kinematics(body)
        if active is None:
            active = body.GetActiveDOFIndices() if body.is_robot else list(range(body.GetDOF()))
        act = np.array(active, dtype=np.int32)
        opt = OrcOptions(self._options, flags, 1 if report else 0, 1 if want_band else 0, band, 0, 0)
        if poses is not None:
            q = np.ascontiguousarray(poses, dtype=np.float64)
            n = q.shape[0]
            opt.pose_mode = 1
        else:
            if bid in self._pose_only:
                raise NotImplementedError("%s has a mimic equation the oracle's FK does not evaluate: pass poses" % body.name)
            q = np.ascontiguousarray(q, dtype=np.float64)
            ncols = len(active) + flatten.affine_count(getattr(body, "affine_dofs", 0))
            q = q.reshape(-1, ncols) if ncols else q.reshape(q.shape[0] if q.ndim == 2 else 1, 0)     # (a body without DOFs: n rows of nothing)
            n = q.shape[0]
        nl = len(body.links) + sum(len(g.body.links) for g in getattr(body, "grabbed", []))   # grabbed links follow the body's own
        verdict = np.zeros(n, dtype=np.uint32)
        fe = np.zeros((n, 3), dtype=np.int32)
        fs = np.zeros((n, 2), dtype=np.int32)
        cnt = np.zeros((n, 2), dtype=np.int64)
        pen = np.zeros(n, dtype=np.float64)
        near = np.zeros(n, dtype=np.uint8)
        cap = 64 * n if report else 0
        ep = np.zeros((max(cap, 1), 4), dtype=np.int32)
        nep = ctypes.c_int64(0)
        fk = np.zeros((n, nl, 8), dtype=np.float64) if want_fk else None
        self.lib.orc_check_batch(self.scene, bid, _ptr(act), len(active), ctypes.byref(opt), _ptr(q), n,
                                 _ptr(verdict), _ptr(fe), _ptr(fs), _ptr(cnt), _ptr(pen), _ptr(near),
                                 _ptr(ep) if report else None, cap, ctypes.addressof(nep), _ptr(fk))
        return dict(verdict=verdict, first_env=fe, first_self=fs, counters=cnt, penetration=pen, near=near,
                    env_pairs=ep[:min(nep.value, cap)] if report else None, fk=fk)

    def distance_batch(self, body, q, active=None, flags=BATCH_ENV | BATCH_SELF, max_distance=0.0, poses=None):
        """Exact minimum separation per configuration (PQP_Distance with both tolerances 0 over every visited pair):
        dict(distance [n], pair [n, 3] {query link, other body, other link})."""
        self._sync()
        bid = self._sync_kinematics(body)
        if active is None:
            active = body.GetActiveDOFIndices() if body.is_robot else list(range(body.GetDOF()))
        act = np.array(active, dtype=np.int32)
        opt = OrcOptions(self._options, flags, 0, 0, 0.0, 0, 0)
        if poses is not None:
            q = np.ascontiguousarray(poses, dtype=np.float64)
            n = q.shape[0]
            opt.pose_mode = 1
        else:
            q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, len(active))
            n = q.shape[0]
        dist = np.zeros(n, dtype=np.float64)
        pair = np.zeros((n, 3), dtype=np.int32)
        self.lib.orc_distance_batch(self.scene, bid, _ptr(act), len(active), ctypes.byref(opt), _ptr(q), n,
                                    float(max_distance), _ptr(dist), _ptr(pair))
        return dict(distance=dist, pair=pair)

    def tri_tri_collision(self, u, v):
        """The reference's TriTriCollision on world-space triangles u, v [3,3] (reference build only): (hit, pos, norm)."""
        u = np.ascontiguousarray(u, dtype=np.float64); v = np.ascontiguousarray(v, dtype=np.float64)
        pos, norm = np.zeros(3), np.zeros(3)
        rc = self.lib.orc_tri_tri_collision(_ptr(u), _ptr(v), _ptr(pos), _ptr(norm))
        if rc < 0:
            raise RuntimeError("TriTriCollision is only available in the reference-linked oracle build")
        return bool(rc), pos, norm

    def pair_queries(self, body, q, triples, active=None, cap=0.0, tolerance=0.0):
        """Per configuration and (query link, other body id, other link) triple: exact distance clipped at `cap`, and
        PQP_Tolerance's verdict at `tolerance`.  Returns (dist [n, m], within [n, m] bool)."""
        self._sync()
        bid = self._sync_kinematics(body)
        if active is None:
            active = body.GetActiveDOFIndices() if body.is_robot else list(range(body.GetDOF()))
        act = np.array(active, dtype=np.int32)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, len(active))
        tr = np.ascontiguousarray(np.array(triples, dtype=np.int32).reshape(-1, 3))
        n, m = q.shape[0], tr.shape[0]
        dist = np.zeros((n, m), dtype=np.float64)
        within = np.zeros((n, m), dtype=np.uint8)
        self.lib.orc_pair_queries(self.scene, bid, _ptr(act), len(act), _ptr(q), n, _ptr(tr), m, float(cap), float(tolerance),
                                  _ptr(dist), _ptr(within))
        return dist, within.astype(bool)

    def pair_band(self, body, q, triples, active=None, band=1e-4):
        """Contact-band class of every (query link, other body id, other link) triple in every configuration, [n, m] uint8:
        0 free beyond the band, 1 free within it, 2 colliding with penetration measure <= band, 3 colliding deeper."""
        self._sync()
        bid = self._sync_kinematics(body)
        if active is None:
            active = body.GetActiveDOFIndices() if body.is_robot else list(range(body.GetDOF()))
        act = np.array(active, dtype=np.int32)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, len(active))
        tr = np.ascontiguousarray(np.array(triples, dtype=np.int32).reshape(-1, 3))
        out = np.zeros((q.shape[0], tr.shape[0]), dtype=np.uint8)
        self.lib.orc_pair_band(self.scene, bid, _ptr(act), len(act), _ptr(q), q.shape[0], _ptr(tr), tr.shape[0], float(band), _ptr(out))
        return out

    def time_batch(self, body, q, active=None, flags=BATCH_ENV | BATCH_SELF):
        self._sync()
        bid = self._sync_kinematics(body)
        if active is None:
            active = body.GetActiveDOFIndices() if body.is_robot else list(range(body.GetDOF()))
        act = np.array(active, dtype=np.int32)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, len(active))
        verdict = np.zeros(q.shape[0], dtype=np.uint32)
        secs = self.lib.orc_time_batch(self.scene, bid, _ptr(act), len(active), self._options, flags, _ptr(q),
                                       q.shape[0], _ptr(verdict))
        return secs, verdict

    def group_checker(self, body, active=None):
        """(function pointer, user block) for cc_planner_set_checker: the reference's serial, early-exit edge loop on this
        oracle.  Keep the returned block alive while the planner uses it; it accumulates configs_checked / seconds."""
        self._sync()
        bid = self._sync_kinematics(body)
        if active is None:
            active = body.GetActiveDOFIndices() if body.is_robot else list(range(body.GetDOF()))
        act = np.array(active, dtype=np.int32)
        blk = OrcGroupChecker(self.scene, act.ctypes.data, bid, len(act), self._options, 0, 0, 0.0)
        blk._keep = act
        return ctypes.cast(self.lib.orc_group_check, ctypes.c_void_p), blk

    def scene_handle(self, body):
        self._sync()
        return self.scene, self._ids[id(body)]

    # ---- single-state queries (bodies' current transforms), pqprave semantics
    def CheckCollision(self, a, b=None, report=None):
        from openrave_b200.kinbody import KinBody, Link
        if report is not None:
            report.Reset(self._options)
        if isinstance(a, KinBody) and b is None:
            use_cb = not (self._options & 0x100) and self.env is not None and self.env.HasRegisteredCollisionCallbacks()
            res = self.check_batch(a, None, flags=BATCH_ENV, report=(report is not None or use_cb),
                                   poses=flatten.transforms8(a.GetQueryLinkTransformations())[None])
            hit = bool(res["verdict"][0] & 1)
            if (report is not None or use_cb) and hit:
                bodies = self._bodies_by_id()
                qlinks = list(a.links) + [l for g in a.grabbed for l in g.body.links]
                found = [(qlinks[l1], bodies[ob].links[l2]) for _, l1, ob, l2 in res["env_pairs"]]
                kept = self.env.filter_collisions_through_callbacks(self._options, found) if use_cb else found   # collisionPQP.h:606-617
                hit = len(kept) > 0
                if report is not None and hit:
                    for l1, l2 in kept:
                        report.AddLinkCollision(l1, l2)
                    report.minDistance = 0
            return hit
        if isinstance(a, Link) and isinstance(b, Link):
            if a.body is b.body:
                hits = self._check_link_pairs_at(a.body, a.body.GetLinkTransformations(),
                                                 [(min(a.index, b.index), max(a.index, b.index))])
                hit = hits[0]
            else:
                hit = self._link_vs_links(a, [b])
            if hit and report is not None:
                report.AddLinkCollision(a, b)
                report.minDistance = 0
            return hit
        if isinstance(a, Link) and isinstance(b, KinBody):
            if b is a.body:
                return False
            hit = False
            for l in b.links:
                if self._link_vs_links(a, [l]):
                    hit = True
                    if report is None:
                        return True
                    report.AddLinkCollision(a, l)
                    report.minDistance = 0
            return hit
        if isinstance(a, Link) and b is None:
            hit = False
            for ob in self.env.GetBodies():
                if ob is a.body:
                    continue
                for l in ob.links:
                    if self._link_vs_links(a, [l]):
                        hit = True
                        if report is None:
                            return True
                        report.AddLinkCollision(a, l)
                        report.minDistance = 0
            return hit
        raise NotImplementedError("CheckCollision overload")

    def CheckCollisionExcluding(self, a, bodyexcluded, linkexcluded, report=None):
        """CheckCollisionP with exclusion lists, literally (collisionPQP.h:432-498): one DoPQP per surviving link pair, in order."""
        from openrave_b200.environment import excluding_pairs
        if report is not None:
            report.Reset(self._options)
        hit = False
        for _, pairs in excluding_pairs(self.env, a, bodyexcluded, linkexcluded):
            found = [(l1, l2) for l1, l2 in pairs if self._link_vs_links(l1, [l2])]
            kept = self.env.filter_collisions_through_callbacks(self._options, found)
            if kept:
                hit = True
                if report is None:
                    return True
                for l1, l2 in kept:
                    report.AddLinkCollision(l1, l2)
                report.minDistance = 0
        return hit

    def _bodies_by_id(self):
        return {self._ids[id(b)]: b for b in self.env.GetBodies() if id(b) in self._ids}

    def _link_vs_links(self, a, others):
        self._sync()
        if not a.enabled:
            return False
        ba = self._ids[id(a.body)]
        Ta = flatten.transforms8([a.t])
        for o in others:
            if not o.enabled:
                continue
            bo = self._ids[id(o.body)]
            if not self.lib.orc_link_valid(self.scene, ba, a.index) or not self.lib.orc_link_valid(self.scene, bo, o.index):
                continue
            To = flatten.transforms8([o.t])
            n = self.lib.orc_collide_pair(self.scene, ba, a.index, _ptr(Ta), bo, o.index, _ptr(To), 0, None, None, 0)
            if n > 0:
                return True
        return False

    def CheckStandaloneSelfCollision(self, body, report=None):
        if len(body.links) <= 1:
            return False
        if report is not None:
            report.Reset(self._options)
        res = self.check_batch(body, None, flags=BATCH_SELF, poses=flatten.transforms8(body.GetQueryLinkTransformations())[None])
        # the standalone check stops at the body's own non-adjacent pairs; a first hit on a grabbed link is not one of them
        l0, l1 = res["first_self"][0]
        hit = bool(res["verdict"][0] & 2) and l0 < len(body.links) and l1 < len(body.links)
        if hit and report is not None:
            report.AddLinkCollision(body.links[l0], body.links[l1])
            report.minDistance = 0
        return hit

    # ---- reference-only extras
    def ref_distance(self, link1, link2, rel_err=200.0, abs_err=0.001):
        """PQP_Distance with pqprave's tolerances (collisionPQP.h:48-49, 630)."""
        self._sync()
        T1, T2 = flatten.transforms8([link1.t]), flatten.transforms8([link2.t])
        return self.lib.orc_ref_distance(self.scene, self._ids[id(link1.body)], link1.index, _ptr(T1),
                                         self._ids[id(link2.body)], link2.index, _ptr(T2), rel_err, abs_err)

    def collide_pair(self, link1, link2, T1=None, T2=None, first_only=False, max_contacts=1 << 16):
        self._sync()
        T1 = flatten.transforms8([link1.t if T1 is None else T1])
        T2 = flatten.transforms8([link2.t if T2 is None else T2])
        cnt = np.zeros(2, dtype=np.int64)
        out = np.zeros((max_contacts, 2), dtype=np.int32)
        n = self.lib.orc_collide_pair(self.scene, self._ids[id(link1.body)], link1.index, _ptr(T1),
                                      self._ids[id(link2.body)], link2.index, _ptr(T2), 1 if first_only else 0,
                                      _ptr(cnt), _ptr(out), max_contacts)
        return int(n), cnt, out[:min(n, max_contacts)]
