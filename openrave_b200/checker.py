"""`cudacollision` — the CollisionCheckerBase-shaped front of libcudacollision.so.

This class is the harness twin of the OpenRAVE plugin in plugins/cudacollision/: same method names
and argument meaning as `CollisionCheckerBase` (include/openrave/collisionchecker.h:186-367) with
pqprave's semantics (plugins/pqprave/collisionPQP.h), plus the batched entry `CheckCollisionBatch`
that `DynamicsCollisionConstraint::Check` (src/libopenrave/planningutils.cpp:3405-3526), the RRT
family (plugins/rplanners/rplanners.h:365,370) and `CacheCollisionChecker` misses
(plugins/configurationcache/cachechecker.cpp:252) call with N DOF vectors at once.

Every collision verdict comes from the CUDA kernels through the C ABI (include/cudacollision.h).
There is no CPU fallback: if the shared library or a CUDA device is missing, construction raises.
"""
import ctypes
import os

import numpy as np

from . import flatten
from .flatten import cc_joint
from .kinbody import KinBody, Link

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUDACOLLISION_LIB") or os.path.join(_HERE, "libcudacollision.so")   # override: A/B builds

CO_Distance, CO_UseTolerance, CO_Contacts, CO_RayAnyHit = 1, 2, 4, 8
CO_ActiveDOFs, CO_AllLinkCollisions, CO_AllGeometryCollisions, CO_AllGeometryContacts, CO_IgnoreCallbacks = \
    0x10, 0x20, 0x40, 0x80, 0x100
BATCH_ENV, BATCH_SELF, BATCH_ALL_PAIRS, BATCH_FIRST_HIT = 1, 2, 4, 8


class CudaCollisionError(RuntimeError):
    pass


class cc_query_info(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("ndof_active", "nlinks", "npairs_env", "npairs_self", "npairs_hoisted",
                                              "mask_words", "total_nodes", "total_tris", "nexternal", "reserved")]


class cc_certificate(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("pair", "col0", "col1", "n0", "n1", "certified_cells", "tightest_cell", "dropped")] + \
               [(n, ctypes.c_double) for n in ("lo0", "width0", "lo1", "width1", "k0", "k1", "delta", "margin", "min_slack",
                                               "min_distance")]


class cc_geom(ctypes.Structure):
    """Mirror of `struct cc_geom` (include/cudacollision.h): one analytic primitive of a link."""
    _fields_ = [("type", ctypes.c_int32), ("reserved", ctypes.c_int32), ("pose", ctypes.c_double * 8), ("dims", ctypes.c_double * 4)]


GEOM_TYPES = {"box": 1, "sphere": 2, "cylinder": 3, "capsule": 4}


class cc_stats(ctypes.Structure):
    _fields_ = [("configs", ctypes.c_uint64), ("bv_tests", ctypes.c_uint64), ("tri_tests", ctypes.c_uint64),
                ("launches", ctypes.c_uint64), ("last_kernel_ms", ctypes.c_double)]


EXPORTS = ["cc_last_error", "cc_version", "cc_scene_create", "cc_scene_destroy", "cc_scene_add_body",
           "cc_scene_remove_body", "cc_body_set_link_mesh", "cc_body_set_link_enabled", "cc_body_set_link_transforms",
           "cc_body_set_attached", "cc_body_set_kinematics", "cc_body_set_dof_values", "cc_body_set_nonadjacent_pairs",
           "cc_scene_commit", "cc_query_create", "cc_query_create_pairs", "cc_query_destroy", "cc_query_get_info",
           "cc_query_get_pairs", "cc_check_batch", "cc_check_batch_device", "cc_check_poses", "cc_fk_batch",
           "cc_scene_synchronize", "cc_query_trace_work", "cc_query_get_work_trace", "cc_scene_set_tri_epsilon", "cc_scene_set_counters", "cc_scene_get_stats",
           "cc_host_alloc", "cc_host_free", "cc_body_set_grabbed", "cc_body_release_grabbed", "cc_query_get_links",
           "cc_distance_batch", "cc_distance_poses", "cc_scene_set_pair_tables", "cc_query_get_table_count",
           "cc_tolerance_batch", "cc_tolerance_poses", "cc_device_fp32_peak", "cc_query_refresh",
           "cc_contacts_poses", "cc_tri_tri_contact", "cc_debug_build_bvh", "cc_debug_broad_overlap", "cc_debug_motion_classes", "cc_body_set_mimic_equation", "cc_query_expand_rows", "cc_debug_mimic_eval", "cc_debug_expand_rows", "cc_debug_fk_program", "cc_query_get_certificates", "cc_debug_pair_rates",
           "cc_scene_trace_timeline", "cc_scene_get_timeline", "cc_scene_create_multi", "cc_scene_get_devices", "cc_query_create_affine",
           "cc_body_set_link_primitives", "cc_scene_set_analytic_primitives"]

_lib = None


def load_library():
    """dlopen libcudacollision.so and declare the prototypes of include/cudacollision.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise CudaCollisionError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    lib.cc_last_error.restype = ctypes.c_char_p
    lib.cc_scene_create.argtypes = [ci, ctypes.POINTER(vp)]
    lib.cc_scene_create_multi.argtypes = [vp, ci, ctypes.POINTER(vp)]
    lib.cc_scene_get_devices.argtypes = [vp, vp, ci, ctypes.POINTER(ci)]
    lib.cc_scene_destroy.argtypes = [vp]
    lib.cc_scene_add_body.argtypes = [vp, ctypes.c_char_p, ci, ci, ctypes.POINTER(ci)]
    lib.cc_scene_remove_body.argtypes = [vp, ci]
    lib.cc_body_set_link_mesh.argtypes = [vp, ci, ci, ctypes.c_char_p, vp, ci, vp, ci]
    lib.cc_body_set_link_enabled.argtypes = [vp, ci, vp]
    lib.cc_body_set_link_primitives.argtypes = [vp, ci, ci, vp, ci]
    lib.cc_scene_set_analytic_primitives.argtypes = [vp, ci]
    lib.cc_body_set_link_transforms.argtypes = [vp, ci, vp]
    lib.cc_body_set_attached.argtypes = [vp, ci, ci, ci]
    lib.cc_body_set_kinematics.argtypes = [vp, ci, ci, ci, vp, vp]
    lib.cc_body_set_grabbed.argtypes = [vp, ci, ci, ci, vp, vp, ci]
    lib.cc_body_release_grabbed.argtypes = [vp, ci, ci]
    lib.cc_query_get_links.argtypes = [vp, vp]
    lib.cc_body_set_dof_values.argtypes = [vp, ci, vp, ci]
    lib.cc_body_set_nonadjacent_pairs.argtypes = [vp, ci, vp, ci]
    lib.cc_scene_commit.argtypes = [vp]
    lib.cc_query_create.argtypes = [vp, ci, vp, ci, ci, ctypes.POINTER(vp)]
    lib.cc_query_create_pairs.argtypes = [vp, ci, vp, ci, ctypes.POINTER(vp)]
    lib.cc_query_create_affine.argtypes = [vp, ci, vp, ci, ci, vp, vp, ci, ctypes.POINTER(vp)]
    lib.cc_query_destroy.argtypes = [vp]
    lib.cc_query_refresh.argtypes = [vp]
    lib.cc_query_get_info.argtypes = [vp, ctypes.POINTER(cc_query_info)]
    lib.cc_query_get_pairs.argtypes = [vp, vp]
    lib.cc_check_batch.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp]
    lib.cc_check_batch_device.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp, vp]
    lib.cc_check_poses.argtypes = [vp, vp, ci, ci, vp, vp, vp, vp]
    lib.cc_fk_batch.argtypes = [vp, vp, ci, vp]
    lib.cc_device_fp32_peak.argtypes = [ci, ctypes.POINTER(cd)]
    lib.cc_contacts_poses.argtypes = [vp, vp, ci, ci, ctypes.POINTER(ci), vp, vp, vp, vp]
    lib.cc_tri_tri_contact.argtypes = [vp, vp, vp, vp]
    lib.cc_tolerance_batch.argtypes = [vp, vp, ci, ci, cd, vp]
    lib.cc_tolerance_poses.argtypes = [vp, vp, ci, ci, cd, vp]
    lib.cc_scene_set_pair_tables.argtypes = [vp, ci]
    lib.cc_query_get_table_count.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(cd)]
    lib.cc_scene_trace_timeline.argtypes = [vp, ci]
    lib.cc_scene_get_timeline.argtypes = [vp, vp, vp, ci, ctypes.POINTER(ci)]
    lib.cc_query_get_certificates.argtypes = [vp, vp, ci, ctypes.POINTER(ci)]
    lib.cc_debug_pair_rates.argtypes = [vp, ci, ci, vp, ci, ci, ci, cd, cd, vp, vp, ctypes.POINTER(ci)]
    lib.cc_debug_motion_classes.argtypes = [vp, ci, ci, ci, vp, vp, ci, ci, vp, ci, vp, vp]
    cpp = ctypes.POINTER(ctypes.c_char_p)
    lib.cc_body_set_mimic_equation.argtypes = [vp, ci, ci, ctypes.c_char_p, cpp, ci]
    lib.cc_query_expand_rows.argtypes = [vp, vp, ctypes.c_int64, vp]
    lib.cc_debug_mimic_eval.argtypes = [ctypes.c_char_p, cpp, ci, vp, ctypes.c_int64, vp, vp]
    lib.cc_debug_expand_rows.argtypes = [vp, ci, ci, vp, cpp, cpp, vp, ci, ci, vp, ctypes.c_int64, vp]
    lib.cc_debug_fk_program.argtypes = [vp, ci, ci, ci, vp, vp, vp, ci, vp, ci, vp, ctypes.c_int64, vp, vp]
    lib.cc_debug_broad_overlap.argtypes = [vp, vp, ci, vp]
    lib.cc_distance_batch.argtypes = [vp, vp, ci, ci, cd, cd, cd, vp, vp, vp]
    lib.cc_distance_poses.argtypes = [vp, vp, ci, ci, cd, cd, cd, vp, vp, vp]
    lib.cc_scene_synchronize.argtypes = [vp]
    lib.cc_query_trace_work.argtypes = [vp, ci]
    lib.cc_query_get_work_trace.argtypes = [vp, vp, ci]
    lib.cc_scene_set_tri_epsilon.argtypes = [vp, cd]
    lib.cc_scene_set_counters.argtypes = [vp, ci]
    lib.cc_scene_get_stats.argtypes = [vp, ctypes.POINTER(cc_stats)]
    lib.cc_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.cc_host_free.argtypes = [vp]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def device_fp32_peak(device=0):
    """Measured FP32 FMA TFLOP/s of a CUDA device (cc_device_fp32_peak)."""
    lib = load_library()
    v = ctypes.c_double(0.0)
    if lib.cc_device_fp32_peak(device, ctypes.byref(v)) != 0:
        raise CudaCollisionError((lib.cc_last_error() or b"").decode())
    return v.value


class PinnedArray(object):
    """A numpy view over a page-locked buffer from cc_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = ctypes.c_void_p()
        if self.lib.cc_host_alloc(max(nbytes, 1), ctypes.byref(p)) != 0:
            raise CudaCollisionError((self.lib.cc_last_error() or b"").decode())
        self.ptr = p
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self.ptr:
                self.array = None
                self.lib.cc_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class BoundQuery(object):
    """A cc_query plus its decoded pair table."""

    def __init__(self, checker, handle, body):
        self.checker = checker
        self.lib = checker.lib
        self.handle = handle
        self.body = body
        info = cc_query_info()
        checker._ck(self.lib.cc_query_get_info(handle, ctypes.byref(info)))
        self.info = info
        self.npairs = info.npairs_env + info.npairs_self
        self.pairs = np.zeros((max(self.npairs, 1), 4), dtype=np.int32)
        checker._ck(self.lib.cc_query_get_pairs(handle, _ptr(self.pairs)))
        self.pairs = self.pairs[:self.npairs]
        # query link index -> (body id, link): the body's own links, then its grabbed bodies' links
        self.links = np.zeros((max(info.nlinks, 1), 2), dtype=np.int32)
        checker._ck(self.lib.cc_query_get_links(handle, _ptr(self.links)))
        self.links = self.links[:info.nlinks]
        self.nlinks_body = len(body.links)
        self.stale_own = False      # the query body itself moved since the last refresh (matters to the FK-mode entries only)
        self.active = []            # DOF index of every joint column (set by bind)
        self.externals = []         # mimic joints whose equation this side evaluates per configuration (CC_MIMIC_EXTERNAL)

    def rows(self, q):
        """The rows the C entry takes: [active joint values | values of the externally evaluated mimic joints | affine values].
        The caller passes [active joint values | affine values]; the middle block is computed here the way SetDOFValues would
        (clamped DOF values in, raw equation values out: the library applies the mimic limit rules)."""
        ncols = max(self.info.ndof_active, 1)
        if not self.externals:
            return np.ascontiguousarray(q, dtype=np.float64).reshape(-1, ncols)
        from .kinbody import external_mimic_values
        ne, na = len(self.externals), len(self.active)
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, ncols - ne)
        body = self.body
        full = np.tile(np.array(body.GetDOFValues(), dtype=np.float64), (q.shape[0], 1))
        full[:, self.active] = q[:, :na]
        for j in body.joints:                                    # SetDOFValues clamps non-circular DOF values first
            if j.dofindex >= 0 and not j.circular:
                full[:, j.dofindex] = np.clip(full[:, j.dofindex], j.lower, j.upper)
        ext = external_mimic_values(body, full, self.externals)
        return np.ascontiguousarray(np.hstack([q[:, :na], ext, q[:, na:]]))

    def refresh(self):
        """cc_query_refresh: re-read captured poses / inactive DOF values after bodies moved."""
        self.checker._ck(self.lib.cc_query_refresh(self.handle))
        self.stale_own = False

    def _fresh_for_fk(self):
        if self.stale_own:
            self.refresh()

    def table_count(self):
        """How many link pairs of this query are covered by a certified free-space table (0 until the tables are built)."""
        n, d, ms = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_double(0.0)
        self.checker._ck(self.lib.cc_query_get_table_count(self.handle, ctypes.byref(n), ctypes.byref(d), ctypes.byref(ms)))
        self.dropped_pairs = d.value
        self.certificate_build_ms = ms.value
        return n.value + d.value

    def certificates(self):
        """cc_query_get_certificates: what the last certificate build proved (a list of cc_certificate records)."""
        n = ctypes.c_int(0)
        self.checker._ck(self.lib.cc_query_get_certificates(self.handle, None, 0, ctypes.byref(n)))
        arr = (cc_certificate * max(1, n.value))()
        self.checker._ck(self.lib.cc_query_get_certificates(self.handle, ctypes.addressof(arr), n.value, ctypes.byref(n)))
        return [arr[i] for i in range(n.value)]

    def link_of(self, qindex):
        """Query link index -> Link object (a grabbed body's link when qindex >= the body's link count)."""
        b, l = self.links[qindex]
        return self.checker.body_from_id(int(b)).links[int(l)]

    def close(self):
        if self.handle:
            self.lib.cc_query_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode_env(self, idx):
        """first_env indices -> [n,3] {query link, other body id, other link}, -1 rows where idx < 0."""
        out = np.full((len(idx), 3), -1, dtype=np.int32)
        ok = idx >= 0
        out[ok] = self.pairs[idx[ok], 0:3]
        return out

    def decode_self(self, idx):
        out = np.full((len(idx), 2), -1, dtype=np.int32)
        ok = idx >= 0
        out[ok, 0] = self.pairs[idx[ok], 0]
        out[ok, 1] = self.pairs[idx[ok], 2]
        return out

    def check(self, q, flags=BATCH_ENV | BATCH_SELF, want_mask=False, out=None):
        """`out` = (verdict u32[n], first_env i32[n], first_self i32[n]) lets a caller reuse (pinned) result buffers."""
        self._fresh_for_fk()
        q = self.rows(q)
        n = q.shape[0]
        if out is not None:
            verdict, fe, fs = out
            assert verdict.dtype == np.uint32 and fe.dtype == np.int32 and fs.dtype == np.int32
            assert len(verdict) >= n and len(fe) >= n and len(fs) >= n
        else:
            verdict = np.zeros(n, dtype=np.uint32)
            fe = np.zeros(n, dtype=np.int32)
            fs = np.zeros(n, dtype=np.int32)
        mask = np.zeros((n, self.info.mask_words), dtype=np.uint32) if want_mask else None
        self.checker._ck(self.lib.cc_check_batch(self.handle, _ptr(q), n, flags, _ptr(verdict), _ptr(fe), _ptr(fs), _ptr(mask)))
        return verdict, fe, fs, mask

    def check_poses(self, T8, flags=BATCH_ENV | BATCH_SELF, want_mask=False):
        T8 = np.ascontiguousarray(T8, dtype=np.float64).reshape(-1, self.info.nlinks, 8)
        n = T8.shape[0]
        verdict = np.zeros(n, dtype=np.uint32)
        fe = np.zeros(n, dtype=np.int32)
        fs = np.zeros(n, dtype=np.int32)
        mask = np.zeros((n, self.info.mask_words), dtype=np.uint32) if want_mask else None
        self.checker._ck(self.lib.cc_check_poses(self.handle, _ptr(T8), n, flags, _ptr(verdict), _ptr(fe), _ptr(fs), _ptr(mask)))
        return verdict, fe, fs, mask

    def distance(self, q=None, flags=BATCH_ENV | BATCH_SELF, max_distance=0.0, poses=None, want_points=True, rel_err=0.0,
                 abs_err=0.0):
        """CO_Distance for n configurations (or n sets of link transforms): (distance [n], pair index [n], points [n, 6])."""
        if poses is not None:
            a = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, self.info.nlinks, 8)
            fn = self.lib.cc_distance_poses
        else:
            self._fresh_for_fk()
            a = self.rows(q)
            fn = self.lib.cc_distance_batch
        n = a.shape[0]
        dist = np.zeros(n, dtype=np.float64)
        pair = np.zeros(n, dtype=np.int32)
        pts = np.zeros((n, 6), dtype=np.float64) if want_points else None
        self.checker._ck(fn(self.handle, _ptr(a), n, flags, float(max_distance), float(rel_err), float(abs_err), _ptr(dist),
                            _ptr(pair), _ptr(pts)))
        return dist, pair, pts

    def within_tolerance(self, tolerance, q=None, flags=BATCH_ENV | BATCH_SELF, poses=None):
        """CO_UseTolerance: [n, npairs] bool, pair p no farther apart than `tolerance` (PQP_Tolerance semantics)."""
        if poses is not None:
            a = np.ascontiguousarray(poses, dtype=np.float64).reshape(-1, self.info.nlinks, 8)
            fn = self.lib.cc_tolerance_poses
        else:
            self._fresh_for_fk()
            a = self.rows(q)
            fn = self.lib.cc_tolerance_batch
        n = a.shape[0]
        mask = np.zeros((n, self.info.mask_words), dtype=np.uint32)
        self.checker._ck(fn(self.handle, _ptr(a), n, flags, float(tolerance), _ptr(mask)))
        bits = (mask[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1
        return bits.reshape(n, -1)[:, :self.npairs].astype(bool)

    def contacts(self, poses, flags=BATCH_ENV | BATCH_SELF, max_contacts=65536):
        """CO_Contacts for one set of link transforms: (pair index [m], triangles [m,2], has_contact [m] bool, pos [m,3],
        norm [m,3], total) — one row per intersecting triangle pair, sorted by (pair, triangle a, triangle b)."""
        T = np.ascontiguousarray(poses, dtype=np.float64).reshape(self.info.nlinks, 8)
        cnt = ctypes.c_int(0)
        pair = np.zeros(max_contacts, dtype=np.int32)
        tris = np.zeros((max_contacts, 2), dtype=np.int32)
        has = np.zeros(max_contacts, dtype=np.uint8)
        pn = np.zeros((max_contacts, 6), dtype=np.float64)
        self.checker._ck(self.lib.cc_contacts_poses(self.handle, _ptr(T), flags, max_contacts, ctypes.byref(cnt), _ptr(pair), _ptr(tris),
                                                    _ptr(has), _ptr(pn)))
        m = min(cnt.value, max_contacts)
        return pair[:m], tris[:m], has[:m].astype(bool), pn[:m, 0:3], pn[:m, 3:6], cnt.value

    def decode_pair(self, idx):
        """pair indices -> [n, 3] {query link, other body id, other link} (oracle layout), -1 rows where idx < 0."""
        out = np.full((len(idx), 3), -1, dtype=np.int32)
        ok = idx >= 0
        out[ok] = self.pairs[idx[ok], 0:3]
        return out

    def fk(self, q):
        self._fresh_for_fk()
        q = self.rows(q)
        n = q.shape[0]
        out = np.zeros((n, self.info.nlinks, 8), dtype=np.float64)
        self.checker._ck(self.lib.cc_fk_batch(self.handle, _ptr(q), n, _ptr(out)))
        return out

    def trace_work(self, enabled=True):
        self.checker._ck(self.lib.cc_query_trace_work(self.handle, 1 if enabled else 0))

    def work_trace(self, n):
        out = np.zeros(n, dtype=np.uint32)
        self.checker._ck(self.lib.cc_query_get_work_trace(self.handle, _ptr(out), n))
        return out

    def check_device(self, d_q_ptr, n, flags, d_verdict_ptr, d_fe_ptr, d_fs_ptr, d_mask_ptr=0, stream=0):
        """Raw device-pointer entry (torch tensors' data_ptr()); asynchronous on `stream`."""
        self._fresh_for_fk()
        self.checker._ck(self.lib.cc_check_batch_device(self.handle, d_q_ptr, n, flags, d_verdict_ptr, d_fe_ptr, d_fs_ptr,
                                                        d_mask_ptr or None, stream or None))


class CudaCollisionChecker(object):
    """RaveCreateCollisionChecker(env, 'cudacollision')."""

    xmlid = "cudacollision"

    def __init__(self, env=None, device=0):
        self.lib = load_library()
        self.env = env
        self.device = device
        self.scene = None
        self._ids = {}
        self._bodies = {}
        self._dirty = set()
        self._kin_dirty = set()
        self._externals = {}
        self._equations_registered = set()
        self._options = 0
        self._queries = {}
        self._epoch = 0
        self._grab_epochs = {}
        self._create_scene()

    # ---- plumbing
    def _ck(self, rc):
        if rc != 0:
            raise CudaCollisionError("cudacollision error %d: %s" % (rc, self.lib.cc_last_error().decode()))

    def _create_scene(self):
        h = ctypes.c_void_p()
        if isinstance(self.device, (list, tuple)):      # "cudacollision 0,1,2,3": one scene replicated on several devices
            devs = (ctypes.c_int * len(self.device))(*[int(d) for d in self.device])
            self._ck(self.lib.cc_scene_create_multi(devs, len(self.device), ctypes.byref(h)))
        else:
            self._ck(self.lib.cc_scene_create(self.device, ctypes.byref(h)))
        self.scene = h

    def __del__(self):
        try:
            self.DestroyEnvironment()
        except Exception:
            pass

    # ---- CollisionCheckerBase surface (collisionchecker.h:200-237)
    def SetCollisionOptions(self, options):
        if options != self._options:
            self._options = int(options)
            self._invalidate()
        return True

    def GetCollisionOptions(self):
        return self._options

    def SetTolerance(self, tol):
        """collisionPQP.h:426-428: enables the tolerance count and sets its threshold."""
        self._tolerance = float(tol)
        self.SetCollisionOptions(self._options | CO_UseTolerance)
        return True

    def InitEnvironment(self):
        self._drop_queries()
        if self.scene is not None:
            self.lib.cc_scene_destroy(self.scene)
            self.scene = None
        self._create_scene()
        if getattr(self, "_analytic", False):
            self._ck(self.lib.cc_scene_set_analytic_primitives(self.scene, 1))
        self._ids, self._bodies, self._dirty, self._kin_dirty = {}, {}, set(), set()
        self._grab_epochs = {}
        self._externals, self._equations_registered = {}, set()
        for b in self.env.GetBodies():
            self.InitKinBody(b)
        return True

    def DestroyEnvironment(self):
        self._drop_queries()
        if self.scene is not None:
            self.lib.cc_scene_destroy(self.scene)
            self.scene = None
        self._ids, self._bodies = {}, {}

    def InitKinBody(self, body):
        if id(body) in self._ids:
            return True
        bid = ctypes.c_int(-1)
        self._ck(self.lib.cc_scene_add_body(self.scene, body.name.encode(), len(body.links), 1 if body.is_robot else 0,
                                            ctypes.byref(bid)))
        self._ids[id(body)] = bid.value
        self._bodies[bid.value] = body
        for l in body.links:
            v, i = flatten.link_mesh_arrays(l)
            self._ck(self.lib.cc_body_set_link_mesh(self.scene, bid.value, l.index, l.name.encode(), _ptr(v), v.shape[0],
                                                    _ptr(i), i.shape[0]))
            prims = getattr(l, "prims", None)
            if prims:       # KinBody::GeometryInfo of a link made of primitives only: used when the analytic mode is on
                arr = (cc_geom * len(prims))()
                for k, g in enumerate(prims):
                    arr[k].type = GEOM_TYPES[g["type"]]
                    for j, x in enumerate(list(g["t"][0]) + list(g["t"][1]) + [0.0]):
                        arr[k].pose[j] = x
                    d = list(g["data"]) + [0.0, 0.0, 0.0, 0.0]
                    for j in range(4):
                        arr[k].dims[j] = d[j]
                self._ck(self.lib.cc_body_set_link_primitives(self.scene, bid.value, l.index, ctypes.addressof(arr), len(prims)))
        self._dirty.add(id(body))
        self._kin_dirty.add(id(body))
        self._invalidate()
        return True

    def RemoveKinBody(self, body):
        bid = self._ids.pop(id(body), None)
        if bid is not None:
            self._bodies.pop(bid, None)
            self._ck(self.lib.cc_scene_remove_body(self.scene, bid))
            self._invalidate()

    def _on_body_changed(self, body, kind="structure"):
        """Pose changes are pushed at the next query and the cached bound queries refreshed (cc_query_refresh); anything that
        can change a pair list (enable flags, grabs, active DOFs, geometry) drops them."""
        self._dirty.add(id(body))
        if kind != "pose":
            self._invalidate()

    def _invalidate(self):
        self._epoch += 1
        self._drop_queries()

    def _drop_queries(self):
        for q in self._queries.values():
            q.close()
        self._queries = {}

    def _sync(self):
        if self.scene is None:
            self.InitEnvironment()
        for b in self.env.GetBodies():
            if id(b) not in self._ids:
                self.InitKinBody(b)
        moved = set()
        for b in self.env.GetBodies():
            if id(b) in self._dirty:
                moved.add(id(b))
                bid = self._ids[id(b)]
                T = flatten.transforms8(b.GetLinkTransformations())
                self._ck(self.lib.cc_body_set_link_transforms(self.scene, bid, _ptr(T)))
                en = np.array([1 if l.enabled else 0 for l in b.links], dtype=np.uint8)
                self._ck(self.lib.cc_body_set_link_enabled(self.scene, bid, _ptr(en)))
                if b.GetDOF() > 0 and id(b) not in self._kin_dirty:
                    cur = np.array(b.GetDOFValues(), dtype=np.float64)
                    self._ck(self.lib.cc_body_set_dof_values(self.scene, bid, _ptr(cur), b.GetDOF()))
        self._dirty.clear()
        for b in self.env.GetBodies():
            ep = getattr(b, "grab_epoch", 0)
            if self._grab_epochs.get(id(b), 0) != ep:
                bid = self._ids[id(b)]
                self._ck(self.lib.cc_body_release_grabbed(self.scene, bid, -1))
                for gid, link, rel, nc in flatten.grab_tables(b, lambda k: self._ids[id(k)]):
                    self._ck(self.lib.cc_body_set_grabbed(self.scene, bid, gid, link, _ptr(rel), _ptr(nc) if len(nc) else None, len(nc)))
                self._grab_epochs[id(b)] = ep
        self._ck(self.lib.cc_scene_commit(self.scene))
        # bound queries captured poses: other bodies moving needs a refresh now, the query body's own motion only matters to
        # the FK-mode entries (base pose, inactive DOF values) and is refreshed lazily there
        for bq in self._queries.values():
            own = {id(bq.body)} | {id(g.body) for g in getattr(bq.body, "grabbed", [])}
            if moved - own:
                bq.refresh()
            elif moved & own:
                bq.stale_own = True

    def _sync_kinematics(self, body, need_pairs=True):
        bid = self._ids[id(body)]
        if id(body) in self._kin_dirty:
            if body.sorted_joints_all:
                ext = []
                arr, n = flatten.joint_table(body, externals=ext)      # mimic equations the device cannot evaluate come back in `ext`
                self._externals[bid] = ext
                aff = flatten.affects_table(body)
                self._ck(self.lib.cc_body_set_kinematics(self.scene, bid, body.GetDOF(), n, ctypes.addressof(arr), _ptr(aff)))
                cur = np.array(body.GetDOFValues(), dtype=np.float64)
                if body.GetDOF() > 0:
                    self._ck(self.lib.cc_body_set_dof_values(self.scene, bid, _ptr(cur), body.GetDOF()))
                if bid in self._equations_registered:        # cc_body_set_kinematics forgets the equations
                    self._send_equations(body, bid, ext)
            self._kin_dirty.discard(id(body))
        if need_pairs and len(body.links) > 1:
            na = np.array(body.GetNonAdjacentLinks(0, checker=self), dtype=np.int32)
            self._ck(self.lib.cc_body_set_nonadjacent_pairs(self.scene, bid, _ptr(na), int(na.size)))
        return bid

    # ---- bound queries
    def bind(self, body, active=None, options=None):
        """cc_query_create for (body, active DOFs, options); cached until the scene changes."""
        self._sync()
        bid = self._sync_kinematics(body)
        for g in getattr(body, "grabbed", []):          # grabbedBody.CheckSelfCollision needs its non-adjacent list
            if len(g.body.links) > 1:
                self._sync_kinematics(g.body)
        opts = self._options if options is None else options
        if active is None and body.is_robot and body.active_dofs is not None:
            active = body.GetActiveDOFIndices()
        affine = int(getattr(body, "affine_dofs", 0)) if body.is_robot else 0
        key = (bid, tuple(active) if active is not None else None, opts, affine, tuple(getattr(body, "affine_axis", ())) if affine else ())
        q = self._queries.get(key)
        if q is None:
            h = ctypes.c_void_p()
            if affine:
                # RobotBase::SetActiveDOFs(indices, affine, axis): the affine values follow the joint values in every row
                act = np.array(list(active) if active is not None else list(range(body.GetDOF())), dtype=np.int32)
                axis = np.array(body.affine_axis, dtype=np.float64)
                qs = np.array(body.affine_quat_start, dtype=np.float64)
                self._ck(self.lib.cc_query_create_affine(self.scene, bid, _ptr(act) if len(act) else None, len(act), affine, _ptr(axis), _ptr(qs),
                                                         opts, ctypes.byref(h)))
            elif active is None:
                self._ck(self.lib.cc_query_create(self.scene, bid, None, 0, opts, ctypes.byref(h)))
            else:
                act = np.array(active, dtype=np.int32)
                self._ck(self.lib.cc_query_create(self.scene, bid, _ptr(act), len(active), opts, ctypes.byref(h)))
            q = BoundQuery(self, h, body)
            q.active = list(active) if active is not None else list(range(body.GetDOF()))
            q.externals = list(self._externals.get(bid, []))
            self._queries[key] = q
        return q

    def _bind_pairs(self, body, triples):
        self._sync()
        bid = self._sync_kinematics(body, need_pairs=False)
        tr = np.ascontiguousarray(np.array(triples, dtype=np.int32).reshape(-1, 3))
        h = ctypes.c_void_p()
        self._ck(self.lib.cc_query_create_pairs(self.scene, bid, _ptr(tr), tr.shape[0], ctypes.byref(h)))
        return BoundQuery(self, h, body)

    def _check_link_pairs_at(self, body, transforms, pairs):
        """CheckCollision(link,link) for many pairs of one body placed at `transforms`
        (KinBody::GetNonAdjacentLinks' initial-pose probe, kinbody.cpp:5563-5577)."""
        if not pairs:
            return []
        bid = self._ids[id(body)] if id(body) in self._ids else None
        if bid is None:
            self.InitKinBody(body)
            bid = self._ids[id(body)]
        q = self._bind_pairs(body, [(a, bid, b) for a, b in pairs])
        try:
            _, _, _, mask = q.check_poses(flatten.transforms8(transforms)[None], BATCH_SELF | BATCH_ALL_PAIRS, want_mask=True)
        finally:
            q.close()
        return [bool((mask[0, p >> 5] >> (p & 31)) & 1) for p in range(len(pairs))]

    # ---- the batched entry
    def CheckCollisionBatch(self, robot, q, flags=BATCH_ENV | BATCH_SELF, active=None, all_pairs=False):
        """N configurations at once.  `q` is [n, nactive] active-DOF values.  Returns a dict:
        verdict[n] (bit0 env, bit1 self), first_env[n,3] {link, other body id, other link},
        first_self[n,2] {link i, link j}, and with all_pairs the pair bitmask + the query (for decoding)."""
        bq = self.bind(robot, active)
        f = flags | (BATCH_ALL_PAIRS if all_pairs else 0)
        verdict, fe, fs, mask = bq.check(q, f, want_mask=all_pairs)
        return dict(verdict=verdict, first_env=bq.decode_env(fe), first_self=bq.decode_self(fs), mask=mask, query=bq,
                    first_env_index=fe, first_self_index=fs)

    def DistanceBatch(self, robot, q, flags=BATCH_ENV | BATCH_SELF, active=None, max_distance=0.0, rel_err=0.0, abs_err=0.0):
        """Minimum separation of N configurations: dict(distance [n], pair [n,3] {query link, other body id, other link},
        points [n,6] world closest points)."""
        bq = self.bind(robot, active)
        dist, pair, pts = bq.distance(q, flags, max_distance, rel_err=rel_err, abs_err=abs_err)
        return dict(distance=dist, pair=bq.decode_pair(pair), pair_index=pair, points=pts, query=bq)

    def _fill_distance(self, bq, body, flags, report):
        """CO_Distance on a single-state query (collisionPQP.h:622-648): minDistance, the closest link pair, one contact."""
        T = flatten.transforms8(body.GetQueryLinkTransformations())[None]
        dist, pair, pts = bq.distance(poses=T, flags=flags)
        if pair[0] >= 0 and dist[0] < report.minDistance:
            report.minDistance = float(dist[0])
            l1, ob, l2, _ = bq.pairs[pair[0]]
            other = bq.link_of(l2) if bq.pairs[pair[0], 3] == 1 else self._bodies[ob].links[l2]
            report.AddLinkCollision(bq.link_of(l1), other)
            p1, p2 = pts[0, 0:3], pts[0, 3:6]
            depth = float(np.linalg.norm(p2 - p1))
            norm = (p2 - p1) / depth if depth > 0 else np.zeros(3)
            report.contacts = [(p1.copy(), norm, -depth)]

    def RegisterMimicEquations(self, body, active=None):
        """Hand the text of the body's non-affine mimic equations to the library (cc_body_set_mimic_equation), so that callers
        without an evaluator of their own — the planner host (cc_planner_set_query) and cc_query_expand_rows — can fill the
        CC_MIMIC_EXTERNAL columns.  This twin keeps evaluating them itself (BoundQuery.rows).  Returns the number registered."""
        bq = self.bind(body, active)                       # makes sure the kinematics went over
        bid = self._ids[id(body)]
        self._send_equations(body, bid, bq.externals)
        self._equations_registered.add(bid)                # sent again whenever the kinematics are
        return len(bq.externals)

    def _send_equations(self, body, bid, externals):
        js = body.sorted_joints_all
        names = (ctypes.c_char_p * max(len(js), 1))(*[j.name.encode() for j in js])
        for e in externals:
            k = [i for i, j in enumerate(js) if j is e][0]
            self._ck(self.lib.cc_body_set_mimic_equation(self.scene, bid, k, e.mimic.text.replace("**", "^").encode(), names, len(js)))

    def ForwardKinematicsBatch(self, robot, q, active=None):
        return self.bind(robot, active).fk(q)

    def body_id(self, body):
        return self._ids[id(body)]

    def body_from_id(self, bid):
        return self._bodies[bid]

    def GetStatistics(self):
        st = cc_stats()
        self._ck(self.lib.cc_scene_get_stats(self.scene, ctypes.byref(st)))
        return dict(configs=st.configs, bv_tests=st.bv_tests, tri_tests=st.tri_tests, launches=st.launches,
                    last_kernel_ms=st.last_kernel_ms)

    def TraceTimeline(self, enabled=True):
        self._ck(self.lib.cc_scene_trace_timeline(self.scene, 1 if enabled else 0))

    def GetTimeline(self):
        """Per-chunk timeline of the last traced batched call: (ms [chunks, 3] {lane free, FK end, traversal end}, lane [chunks])."""
        n = ctypes.c_int(0)
        self._ck(self.lib.cc_scene_get_timeline(self.scene, None, None, 0, ctypes.byref(n)))
        ms = np.zeros((max(n.value, 1), 3), dtype=np.float64)
        lane = np.zeros(max(n.value, 1), dtype=np.int32)
        self._ck(self.lib.cc_scene_get_timeline(self.scene, _ptr(ms), _ptr(lane), n.value, ctypes.byref(n)))
        return ms[:n.value], lane[:n.value]

    def SetCounters(self, enabled):
        self._ck(self.lib.cc_scene_set_counters(self.scene, 1 if enabled else 0))

    def SetPairTables(self, enabled):
        self._ck(self.lib.cc_scene_set_pair_tables(self.scene, 1 if enabled else 0))
        self._invalidate()

    def SetAnalyticPrimitives(self, enabled):
        """fclrave semantics: links made of box / sphere / cylinder primitives collide as the exact shapes (cc_analytic.cuh)."""
        self._ck(self.lib.cc_scene_set_analytic_primitives(self.scene, 1 if enabled else 0))
        self._analytic = bool(enabled)
        self._invalidate()

    def SetTriangleEpsilon(self, eps):
        self._ck(self.lib.cc_scene_set_tri_epsilon(self.scene, float(eps)))

    # ---- single-state overloads: the bodies' CURRENT link transforms, no FK (collisionPQP.h:172-329)
    def CheckCollision(self, a, b=None, report=None):
        if report is not None:
            report.Reset(self._options)
        if isinstance(a, KinBody) and b is None:
            bq = self.bind(a)
            T = flatten.transforms8(a.GetQueryLinkTransformations())[None]
            # collision callbacks may veto a colliding pair (collisionPQP.h:569-617): then every colliding pair is needed
            use_cb = not (self._options & CO_IgnoreCallbacks) and self.env is not None and self.env.HasRegisteredCollisionCallbacks()
            want_all = report is not None or use_cb
            verdict, fe, _, mask = bq.check_poses(T, BATCH_ENV | (BATCH_ALL_PAIRS if want_all else 0), want_mask=want_all)
            hit = bool(verdict[0] & 1)
            vetoed = set()
            if hit and want_all:
                found = []
                for p in range(bq.info.npairs_env):
                    if (mask[0, p >> 5] >> (p & 31)) & 1:
                        l1, ob, l2, _ = bq.pairs[p]
                        found.append((bq.link_of(l1), self._bodies[ob].links[l2]))
                kept = self.env.filter_collisions_through_callbacks(self._options, found) if use_cb else found
                vetoed = set((id(a_), id(b_)) for a_, b_ in found) - set((id(a_), id(b_)) for a_, b_ in kept)
                hit = len(kept) > 0
                if hit and report is not None:
                    for l1, l2 in kept:
                        report.AddLinkCollision(l1, l2)
                    report.minDistance = 0
            if hit and report is not None and (self._options & CO_Contacts):
                # one CONTACT per intersecting triangle pair that TriTriCollision accepts (collisionPQP.h:591-603), depth 0
                pidx, _, has, pos, norm, _ = bq.contacts(T, BATCH_ENV)
                for p, ok, x, nrm in zip(pidx, has, pos, norm):
                    if ok:
                        l1, ob, l2, _ = bq.pairs[p]
                        key = (bq.link_of(l1), self._bodies[ob].links[l2])
                        if (id(key[0]), id(key[1])) in vetoed:
                            continue
                        report.contacts_by_pair.setdefault(key, []).append((x.copy(), nrm.copy(), 0.0))
            if report is not None and (self._options & CO_Distance) and not hit:
                self._fill_distance(bq, a, BATCH_ENV, report)
            if report is not None and (self._options & CO_UseTolerance):
                # numWithinTol = other bodies with at least one link pair within tolerance (collisionPQP.h:443-497)
                w = bq.within_tolerance(getattr(self, "_tolerance", 0.0), flags=BATCH_ENV, poses=T)[0]
                report.numWithinTol = len(set(int(bq.pairs[p, 1]) for p in np.nonzero(w[:bq.info.npairs_env])[0]))
            return hit
        if isinstance(a, KinBody) and isinstance(b, KinBody):
            triples = [(l1.index, self.body_id(b), l2.index) for l1 in a.links for l2 in b.links]
            return self._single_pairs(a, triples, report)
        if isinstance(a, Link) and isinstance(b, Link):
            if a.body is b.body:
                hit = self._check_link_pairs_at(a.body, a.body.GetLinkTransformations(),
                                                [(a.index, b.index)])[0] if a.enabled and b.enabled else False
                if hit and report is not None:
                    report.AddLinkCollision(a, b)
                    report.minDistance = 0
                return hit
            self._sync()
            return self._single_pairs(a.body, [(a.index, self.body_id(b.body), b.index)], report)
        if isinstance(a, Link) and isinstance(b, KinBody):
            if b is a.body:
                return False
            self._sync()
            return self._single_pairs(a.body, [(a.index, self.body_id(b), l.index) for l in b.links], report)
        if isinstance(a, Link) and b is None:
            self._sync()
            triples = [(a.index, self.body_id(ob), l.index) for ob in self.env.GetBodies() if ob is not a.body for l in ob.links]
            return self._single_pairs(a.body, triples, report)
        raise CudaCollisionError("CheckCollision overload not supported (rays throw in pqprave as well)")

    def CheckCollisionExcluding(self, a, bodyexcluded, linkexcluded, report=None):
        """CheckCollision(body | link, vbodyexcluded, vlinkexcluded, report) (collisionchecker.h:249-263; pqprave
        collisionPQP.h:314-329, 432-498): the surviving link pairs go through one explicit-pair query per attached body."""
        from .environment import excluding_pairs
        if report is not None:
            report.Reset(self._options)
        self._sync()
        hit = False
        sub = type(report)() if report is not None else None
        for qb, pairs in excluding_pairs(self.env, a, bodyexcluded, linkexcluded):
            triples = [(l1.index, self.body_id(l2.body), l2.index) for l1, l2 in pairs]
            if self._single_pairs(qb, triples, sub):
                hit = True
                if report is None:
                    return True
                for l1, l2 in sub.vLinkColliding:
                    report.AddLinkCollision(l1, l2)
                report.minDistance = 0
            if sub is not None:
                sub.Reset(self._options)
        return hit

    def _single_pairs(self, body, triples, report):
        if not triples:
            return False
        q = self._bind_pairs(body, triples)
        try:
            T = flatten.transforms8(body.GetLinkTransformations())[None]
            verdict, _, _, mask = q.check_poses(T, BATCH_ENV | BATCH_SELF | BATCH_ALL_PAIRS, want_mask=True)
            hit = bool(verdict[0])
            if hit:
                found = []
                for p in range(q.npairs):
                    if (mask[0, p >> 5] >> (p & 31)) & 1:
                        l1, ob, l2, _ = q.pairs[p]
                        found.append((body.links[l1], self._bodies[ob].links[l2]))
                kept = self.env.filter_collisions_through_callbacks(self._options, found) if self.env is not None else found
                hit = len(kept) > 0
                if hit and report is not None:
                    for l1, l2 in kept:
                        report.AddLinkCollision(l1, l2)
                    report.minDistance = 0
        finally:
            q.close()
        return hit

    def CheckStandaloneSelfCollision(self, body, report=None):
        """collisionPQP.h:364-386: the body's non-adjacent list in order, first hit wins."""
        if len(body.links) <= 1:
            return False
        if report is not None:
            report.Reset(self._options)
        bq = self.bind(body)
        T = flatten.transforms8(body.GetQueryLinkTransformations())[None]
        verdict, _, fs, _ = bq.check_poses(T, BATCH_SELF)
        # the body's own non-adjacent pairs come first in the pair order: a first hit among them is the standalone verdict
        # (pairs that involve grabbed links belong to KinBody::CheckSelfCollision, not to this call)
        hit = False
        if verdict[0] & 2:
            l0, l1 = bq.decode_self(fs)[0]
            hit = l0 < bq.nlinks_body and l1 < bq.nlinks_body
            if hit and report is not None:
                report.AddLinkCollision(body.links[l0], body.links[l1])
                report.minDistance = 0
        if report is not None and (self._options & CO_Distance) and not hit and not getattr(body, "grabbed", []):
            self._fill_distance(bq, body, BATCH_SELF, report)
        return hit


def RaveCreateCollisionChecker(env, name):
    """Name-based factory like RaveCreateCollisionChecker(env, name): 'cudacollision [device]'."""
    parts = name.split()
    if parts and parts[0].lower() == "cudacollision":
        if len(parts) > 1 and "," in parts[1]:
            return CudaCollisionChecker(env, [int(x) for x in parts[1].split(",") if x != ""])
        dev = int(parts[1]) if len(parts) > 1 else 0
        return CudaCollisionChecker(env, dev)
    return None
