/* cudacollision.h — C ABI of the B200 (sm_100a) batched collision checker.
 *
 * This is the drop-in boundary: an OpenRAVE `CollisionCheckerBase` plugin
 * (plugins/cudacollision/, loaded through RaveCreateCollisionChecker / SetCollisionChecker) keeps the
 * C++ object model on its side and drives the GPU through these entry points only.  Plain pointers
 * and sizes, no C++ types, no exceptions: every function returns a cc_status and
 * cc_last_error() describes the last failure on the calling thread.
 *
 * Reference interfaces replaced (paths relative to the OpenRAVE tree):
 *   include/openrave/collisionchecker.h:186-367   class CollisionCheckerBase
 *   plugins/pqprave/collisionPQP.h                the pqp back-end whose semantics are kept
 *   src/libopenrave/kinbody.cpp:2121-2485         KinBody::SetDOFValues (batched FK on the device)
 *   src/libopenrave/planningutils.cpp:3405-3526   the per-sample loop the batched entry replaces
 *
 * Conventions
 *   transform  = 8 doubles  {qw, qx, qy, qz, tx, ty, tz, 0}   (RaveTransform: scalar-first quaternion,
 *                include/openrave/geometry.h:395-525)
 *   link pair  = packed int  i | (j << 16), i < j             (src/libopenrave/kinbody.cpp:5512-5580)
 *   Ownership: the caller owns every host buffer it passes; the library owns all device memory.
 *   Threading: one cc_scene per thread (mirrors "one checker per environment, caller holds the
 *   environment lock", collisionchecker.h:183, environment-core.h:1493).
 */
#ifndef CUDACOLLISION_H
#define CUDACOLLISION_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cc_scene cc_scene;   /* one per CollisionCheckerBase instance */
typedef struct cc_query cc_query;   /* a bound batched query: (body, active DOFs, options) */

enum cc_status {
    CC_OK = 0,
    CC_ERR_INVALID = 1,      /* bad argument */
    CC_ERR_CUDA = 2,         /* CUDA runtime failure (message holds cudaGetErrorString) */
    CC_ERR_STATE = 3,        /* call out of sequence (e.g. query on an uncommitted scene) */
    CC_ERR_UNSUPPORTED = 4,  /* e.g. ray queries: pqprave throws too (collisionPQP.h:331-356) */
    CC_ERR_NOMEM = 5
};

/* Same bit values as OpenRAVE::CollisionOptions (collisionchecker.h:30-49). */
enum cc_collision_options {
    CC_CO_Distance = 1,
    CC_CO_UseTolerance = 2,
    CC_CO_Contacts = 4,
    CC_CO_RayAnyHit = 8,
    CC_CO_ActiveDOFs = 0x10,
    CC_CO_AllLinkCollisions = 0x20,
    CC_CO_AllGeometryCollisions = 0x40,
    CC_CO_AllGeometryContacts = 0x80,
    CC_CO_IgnoreCallbacks = 0x100
};

/* Joint kinds of the FK chain program (KinBody::JointType, 1-DOF joints). */
enum cc_joint_type { CC_JOINT_REVOLUTE = 1, CC_JOINT_PRISMATIC = 2 };

/* One step of KinBody::SetDOFValues' joint loop (kinbody.cpp:2286-2481), listed in
 * GetDependencyOrderedJointsAll() order:  T_child = T_parent * (left * J(value) * right).
 * `left`/`right` are Joint::GetInternalHierarchyLeftTransform/RightTransform (kinbodyjoint.cpp:1567-1582),
 * `axis` is GetInternalHierarchyAxis(0).  A static joint has is_static=1 and right = identity
 * (kinbodyjoint.cpp:1527-1551).  value = q[dof_index], or mimic_a*q[mimic_dof]+mimic_b for an
 * affine mimic joint (dof_index = -1), clamped to [lower, upper] unless circular. */
typedef struct cc_joint {
    int32_t type;          /* cc_joint_type */
    int32_t parent_link;   /* Joint::GetHierarchyParentLink()->GetIndex() */
    int32_t child_link;    /* Joint::GetHierarchyChildLink()->GetIndex() */
    int32_t dof_index;     /* >= 0 for an active joint, -1 for passive/mimic */
    int32_t mimic_dof;     /* DOF the mimic equation reads, -1 if not mimic, CC_MIMIC_EXTERNAL: evaluated by the caller (below) */
    int32_t is_static;     /* Joint::IsStatic() */
    int32_t is_circular;   /* Joint::IsCircular(0) */
    int32_t reserved;      /* CC_MIMIC_EXTERNAL: ordinal (0, 1, ...) of this joint among the body's externally evaluated ones */
    double mimic_a, mimic_b;
    double lower, upper;   /* Joint::GetLimits */
    double left[8];
    double right[8];
    double axis[4];        /* xyz + pad */
} cc_joint;

/* A mimic joint whose equation is not a*q[dof]+b (an fparser expression, several source joints): the reference evaluates it on
 * the host per SetDOFValues call (kinbody.cpp:2308-2320, Joint::_Eval) and so does the caller here — for every configuration of a
 * batch its raw value goes into column (number of active joint DOFs + ordinal) of q, before the affine columns; the library
 * applies the mimic limit rules (kinbody.cpp:2326-2386) and the joint transform.  Pairs whose relative pose such a joint can
 * change are never hoisted or certified. */
enum { CC_MIMIC_EXTERNAL = -2 };

/* Flags of the batched entry. */
enum cc_batch_flags {
    CC_BATCH_ENV = 1,        /* Environment::CheckCollision(robot) */
    CC_BATCH_SELF = 2,       /* KinBody::CheckSelfCollision(robot) */
    CC_BATCH_ALL_PAIRS = 4,  /* no early exit: fill the complete per-pair bitmask (report != NULL path) */
    CC_BATCH_FIRST_HIT = 8   /* verdict only, the report == NULL path of the reference: CheckCollision(robot) returns at its first
                              * colliding pair (collisionPQP.h:288-290) and the planners' _CheckState skips the self check once the
                              * environment check failed (planningutils.cpp:2523-2548).  CC_VERDICT_ENV is exact; CC_VERDICT_SELF is
                              * exact when CC_VERDICT_ENV is clear and otherwise only set if a self contact happened to be found;
                              * first_env / first_self come back as -1.  Ignored together with CC_BATCH_ALL_PAIRS. */
};

/* Verdict bits written per configuration. */
enum cc_verdict_bits { CC_VERDICT_ENV = 1, CC_VERDICT_SELF = 2 };

typedef struct cc_query_info {
    int32_t ndof_active;       /* columns of q */
    int32_t nlinks;            /* links of the query body */
    int32_t npairs_env;        /* env pairs in reference order: indices [0, npairs_env) */
    int32_t npairs_self;       /* self pairs in reference order: indices [npairs_env, npairs_env+npairs_self) */
    int32_t npairs_hoisted;    /* pairs whose verdict cannot change with the active DOFs (evaluated once at bind) */
    int32_t mask_words;        /* uint32 words per configuration in the pair bitmask */
    int32_t total_nodes;       /* BVH nodes resident on the device */
    int32_t total_tris;        /* triangles resident on the device */
    int32_t nexternal;         /* of ndof_active: columns holding CC_MIMIC_EXTERNAL joint values (after the active joint DOFs) */
    int32_t reserved;
} cc_query_info;

typedef struct cc_stats {
    uint64_t configs;          /* configurations processed since creation */
    uint64_t bv_tests;         /* box-box tests executed on the device (when counters are enabled) */
    uint64_t tri_tests;        /* triangle-triangle tests executed on the device */
    uint64_t launches;         /* kernels launched by this library */
    double   last_kernel_ms;   /* device time of the most recent batch kernel (CUDA events) */
} cc_stats;

const char* cc_last_error(void);
int cc_version(void);

/* ---- scene lifecycle: CollisionCheckerBase ctor / InitEnvironment / DestroyEnvironment (collisionchecker.h:228-229) */
int cc_scene_create(int cuda_device, cc_scene** out);
/* One scene replicated on several devices of the box (north_star: "batches shard across the 8 GPUs of one box with the scene
 * replicated per GPU"; the reference's idiom is one cloned environment per host thread, src/cppexamples/
 * ormultithreadedplanning.cpp:27-33, src/libopenrave-core/environment-core.h:3657-3662).  The returned handle is used like
 * any other scene: every scene / body / query call is mirrored onto the replicas (the BVHs are built once and handed over),
 * and the host-buffer batch entries — cc_check_batch, cc_fk_batch, cc_distance_batch — cut a batch of >= 2 x 16384
 * configurations into contiguous ranges, one per device, each device copying ITS slice of the caller's buffers in and its
 * results out (no collective: the consumer is the host; page-locked buffers from cc_host_alloc are portable across the
 * devices).  Results are bit-identical to the single-device scene's.  Smaller batches, the single-state entries
 * (cc_check_poses, cc_contacts_poses, ...) and cc_check_batch_device run on cuda_devices[0] only.  A device may be listed
 * more than once (the replicas then share it: that is how a one-GPU box tests the sharding). */
int cc_scene_create_multi(const int* cuda_devices, int ndevices, cc_scene** out);
int cc_scene_get_devices(const cc_scene* s, int* out_devices, int max_devices, int* out_count);
int cc_scene_destroy(cc_scene* s);

/* ---- InitKinBody / RemoveKinBody (collisionchecker.h:233-237; pqprave: collisionPQP.h:89-122) */
int cc_scene_add_body(cc_scene* s, const char* name, int nlinks, int is_robot, int* out_body);
int cc_scene_remove_body(cc_scene* s, int body);
/* Link::GetCollisionData(): link-frame vertices (xyz doubles) + triangle indices (TriMesh, openrave.h:2501-2505) */
int cc_body_set_link_mesh(cc_scene* s, int body, int link, const char* link_name,
                          const double* vertices_xyz, int nvertices, const int32_t* triangles, int ntriangles);
/* ---- analytic primitives (fclrave semantics, off by default).  fclrave gives every box / sphere / cylinder / capsule geometry its
 * exact convex shape (plugins/fclrave/fclspace.cpp:475-602) where pqprave, the primary reference, sees the tessellated triangle
 * soup of the whole link (Link::GetCollisionData).  cc_body_set_link_primitives hands a link's primitive geometries over
 * (KinBody::GeometryInfo: type, pose in the link frame, _vGeomData: box half extents; sphere radius; cylinder / capsule radius
 * and height along the geometry's z axis); the link's tessellated mesh is still set with cc_body_set_link_mesh.  After
 * cc_scene_set_analytic_primitives(scene, 1) every link pair with such a link is decided on the exact shapes — primitive x
 * primitive by boolean GJK, primitive x mesh by GJK against the triangles its bounding box reaches — in queries bound
 * afterwards.  FCL itself is an external, un-vendored fork: "parity unpinned"; the exact shapes contain their tessellation, so
 * this mode reports a superset of the tessellated contacts, the difference being pairs closer than the tessellation's sagitta. */
enum cc_geom_type { CC_GEOM_BOX = 1, CC_GEOM_SPHERE = 2, CC_GEOM_CYLINDER = 3, CC_GEOM_CAPSULE = 4 };
typedef struct cc_geom {
    int32_t type;          /* cc_geom_type */
    int32_t reserved;
    double pose[8];        /* geometry pose in the link frame, {qw qx qy qz tx ty tz 0} */
    double dims[4];        /* box: half extents x y z; sphere: radius; cylinder, capsule: radius, height */
} cc_geom;
int cc_body_set_link_primitives(cc_scene* s, int body, int link, const cc_geom* geoms, int ngeoms);
int cc_scene_set_analytic_primitives(cc_scene* s, int enabled);
/* Link::IsEnabled() for every link */
int cc_body_set_link_enabled(cc_scene* s, int body, const uint8_t* enabled);
/* KinBody::GetLinkTransformations(): [nlinks][8] */
int cc_body_set_link_transforms(cc_scene* s, int body, const double* transforms8);
/* KinBody::IsAttached (grabbed / attached bodies are never checked against each other) */
int cc_body_set_attached(cc_scene* s, int body_a, int body_b, int attached);

/* KinBody::Grab / Release / ReleaseAllGrabbed (src/libopenrave/kinbodygrab.cpp).  From now on the links of `grabbed_body`
 * ride on `link` of `body`: T(grabbed link k) = T(link) * rel_transforms8[k] (KinBody::_UpdateGrabbedBodies), in queries
 * bound on `body` AFTER this call:
 *   - they are query links too: cc_query_info.nlinks counts them after the body's own links, cc_check_poses / cc_fk_batch
 *     rows carry them, cc_query_get_links names them;
 *   - env pairs: the grabbed body's links against every non-attached body (collisionPQP.h:314-329 walks all attached bodies);
 *   - self pairs (KinBody::CheckSelfCollision, kinbodycollision.cpp:83-205): every link listed in `noncolliding_links`
 *     ((body, link) pairs: Grabbed::_listNonCollidingLinksWhenGrabbed, i.e. links of `body` and of its other grabbed bodies
 *     that did not touch the grabbed body at grab time) x the grabbed body's links, the grabbed body's own non-adjacent
 *     pairs, and grabbed-vs-grabbed pairs.
 * Also marks the two bodies attached.  cc_body_release_grabbed with grabbed_body < 0 releases everything. */
int cc_body_set_grabbed(cc_scene* s, int body, int grabbed_body, int link, const double* rel_transforms8,
                        const int32_t* noncolliding_links, int n_noncolliding);
int cc_body_release_grabbed(cc_scene* s, int body, int grabbed_body);

/* ---- kinematics of a body, read off a computed KinBody (used by the batched entry only) */
int cc_body_set_kinematics(cc_scene* s, int body, int ndof, int njoints, const cc_joint* joints_in_fk_order,
                           const uint8_t* dof_affects_link /* [ndof][nlinks], KinBody::DoesAffect != 0 */);

/* Optional evaluator for CC_MIMIC_EXTERNAL joints, for callers that have the equation text but no evaluator of their own (from
 * outside KinBody the reference offers none: OpenRAVEFunctionParserReal is opaque, kinbody.h:27-28).  `equation` is the joint's
 * mimic_pos string (Joint::GetMimicEquation(0, 0)) in fparser syntax over joint names; joint_names[njoints] names the joints of
 * cc_body_set_kinematics in the same order.  The library compiles it (operators () ^ unary- ! * / % + - = != < <= > >= & |,
 * fparser's elementary functions, if(), OpenRAVE's SSSA / SASA / SASS; inline variables and the polyroots* helpers are refused with
 * CC_ERR_UNSUPPORTED and the position of the error in cc_last_error()).  Call after cc_body_set_kinematics, which forgets all
 * equations. */
int cc_body_set_mimic_equation(cc_scene* s, int body, int joint, const char* equation, const char* const* joint_names, int njoints);
/* KinBody::GetDOFValues(): the values the DOFs that are NOT active in a query keep (robot.cpp:1193-1225) */
int cc_body_set_dof_values(cc_scene* s, int body, const double* values, int ndof);
/* KinBody::GetNonAdjacentLinks(0): packed pairs in CompareNonAdjacentFarthest order */
int cc_body_set_nonadjacent_pairs(cc_scene* s, int body, const int32_t* packed_pairs, int npairs);

/* Build/refresh the BVHs of new or changed links and upload them (idempotent). */
int cc_scene_commit(cc_scene* s);

/* ---- bound queries
 * cc_query_create: pairs follow pqprave's iteration order — env: every other (non-attached, non-excluded)
 *   body in environment order x query link x other link (collisionPQP.h:432-498); self: the body's
 *   non-adjacent list (collisionPQP.h:364-386); options: CC_CO_ActiveDOFs filters links and pairs
 *   (collisionPQP.h:710-735, robot.cpp:2023-2084).
 * cc_query_create_pairs: explicit list of (query link, other body, other link) triples — the
 *   CheckCollision(link,link) / CheckCollision(link,body) overloads and KinBody::GetNonAdjacentLinks'
 *   initial-pose probe (kinbody.cpp:5563-5577). */
int cc_query_create(cc_scene* s, int body, const int32_t* active_dofs, int nactive, int options, cc_query** out);
int cc_query_create_pairs(cc_scene* s, int body, const int32_t* triples, int ntriples, cc_query** out);
/* cc_query_create with AFFINE active DOFs (RobotBase::SetActiveDOFs(dofindices, affine, rotationaxis); robot.cpp:1193-1225,
 * planner.cpp:658-667 — base planning).  affine_dofs: OpenRAVE::DOFAffine bits (same values).  A row of q then holds the nactive
 * joint values followed by X | Y | Z (those that are set) and ONE rotation: RotationAxis (angle about rotation_axis3),
 * Rotation3D (axis * angle, 3 values) or RotationQuat (4 values; normalised and left-multiplied by quat_limit_start4 =
 * RobotBase::_vRotationQuatLimitStart, NULL = identity): RaveGetTransformFromAffineDOFValues (libopenrave.cpp:1868-1924)
 * applied to the body's pose at bind time (the components that are not active keep their bound values).  Every link moves
 * with the base, so no environment pair is hoisted; self pairs are hoisted as usual.  No free-space certificates. */
enum cc_affine_dofs { CC_DOF_X = 1, CC_DOF_Y = 2, CC_DOF_Z = 4, CC_DOF_RotationAxis = 8, CC_DOF_Rotation3D = 16, CC_DOF_RotationQuat = 32 };
int cc_query_create_affine(cc_scene* s, int body, const int32_t* active_dofs, int nactive, int affine_dofs, const double* rotation_axis3,
                           const double* quat_limit_start4, int options, cc_query** out);
/* A bound query captures, at creation, the poses of every link it does not move itself (other bodies' links, links of the
 * query body no active DOF moves, the base link) and the values of the inactive DOFs.  After cc_body_set_link_transforms /
 * cc_body_set_dof_values on ANY body, cc_query_refresh re-reads them (a few small uploads and the re-evaluation of the
 * pose-invariant pairs; certificates are rebuilt later) — the pair lists, hence link enable flags, attachments, grabs and
 * meshes, must be unchanged, otherwise create a new query.  This is what lets the single-state overloads follow moving
 * bodies the way pqprave does (it reads every transform per call, collisionPQP.h:438-463) without re-binding. */
int cc_query_refresh(cc_query* q);
int cc_query_destroy(cc_query* q);
int cc_query_get_info(const cc_query* q, cc_query_info* info);

/* Host-side: rows [active joint DOFs | affine DOFs] -> the rows the batched entries take, [active joint DOFs | values of the
 * CC_MIMIC_EXTERNAL joints | affine DOFs] (cc_query_info.ndof_active columns), using the equations registered with
 * cc_body_set_mimic_equation: per row the value part of SetDOFValues' joint loop (kinbody.cpp:2149-2224, 2286-2400) — DOF values
 * clamped, inactive DOFs at the body's current values, joints in the listed order, an equation reading a passive joint gets the
 * value the limit rules gave it.  An equation that fails to evaluate (fparser's EvalError: division by zero, sqrt / log / trig
 * domain) fails the call with CC_ERR_INVALID naming the joint and the row; the reference logs a warning and keeps the previous
 * passive value (kinbody.cpp:2318-2320).  CC_ERR_STATE when an external joint has no equation. */
int cc_query_expand_rows(cc_query* q, const double* q_in, int64_t n, double* q_out);
/* pair index -> {query link, other body, other link, 0=env|1=self}, [npairs][4] */
int cc_query_get_pairs(const cc_query* q, int32_t* out_quads);
/* query link index -> {body, link}, [nlinks][2]: the body's own links first, then the links of its grabbed bodies */
int cc_query_get_links(const cc_query* q, int32_t* out_pairs);

/* ---- the batched entry (new; SURVEY.md §8b "Batched entry").  q is row-major [n][ndof_active] doubles of
 * ACTIVE DOF values (RobotBase::SetActiveDOFValues semantics, robot.cpp:1193-1225); the other DOFs keep
 * the values the body had at cc_query_create.  Outputs (any may be NULL):
 *   verdict[n]        cc_verdict_bits
 *   first_env[n]      index of the first colliding env pair in reference order, -1 if none
 *   first_self[n]     index of the first colliding self pair (global pair index), -1 if none
 *   pair_mask[n][mask_words]  bit p set <=> pair p collides (complete only with CC_BATCH_ALL_PAIRS)
 * Host-buffer variant: copies in, runs, copies out, returns when done. */
int cc_check_batch(cc_query* q, const double* dof_values, int n, int flags,
                   uint32_t* verdict, int32_t* first_env, int32_t* first_self, uint32_t* pair_mask);
/* Device-buffer variant: all pointers are device memory on the scene's device, work is enqueued on
 * `cuda_stream` (a cudaStream_t, NULL = legacy default stream) and NOT synchronised.
 * ONE call in flight per scene: the traversal's spill areas, work counters and per-chunk pose scratch belong to the scene and
 * its queries, not to the stream.  Enqueue the next call on the SAME stream (stream order is enough), or synchronise first;
 * two calls on different streams of one scene would share that state.  (The call that first crosses 16k configurations
 * also builds the free-space certificates synchronously, see cc_scene_set_pair_tables.) */
int cc_check_batch_device(cc_query* q, const double* d_dof_values, int n, int flags,
                          uint32_t* d_verdict, int32_t* d_first_env, int32_t* d_first_self,
                          uint32_t* d_pair_mask, void* cuda_stream);
/* Single-state flavour: link transforms given instead of DOF values, [n][nlinks][8] doubles
 * (CheckCollision overloads read the bodies' CURRENT transforms and never run FK, collisionPQP.h:438-447). */
int cc_check_poses(cc_query* q, const double* link_transforms8, int n, int flags,
                   uint32_t* verdict, int32_t* first_env, int32_t* first_self, uint32_t* pair_mask);
/* Batched KinBody::SetDOFValues: out [n][nlinks][8] doubles (validation of the 1e-6 m / 1e-6 rad bound). */
int cc_fk_batch(cc_query* q, const double* dof_values, int n, double* out_transforms8);

/* ---- CO_Distance (collisionPQP.h:622-648: PQP_Distance on every link pair the query visits, CollisionReport::minDistance
 * = the smallest, with the two closest points as contact 0 of that pair).  For every configuration: the minimum separation
 * over the query's env pairs (CC_BATCH_ENV) and self pairs (CC_BATCH_SELF), 0 when some pair touches or penetrates.
 *   max_distance      search radius in metres (pairs farther apart are not refined; the result is then max_distance),
 *                     <= 0: unbounded
 *   rel_err, abs_err  PQP_Distance's tolerances (pqp/PQP.cpp:717-721): 0, 0 = exact; with them the result r obeys
 *                     exact <= r <= max(exact + abs_err, exact * (1 + rel_err)) and the search visits far fewer boxes
 *                     (pqprave itself runs with 200 and 0.001, collisionPQP.h:48-49)
 *   min_distance[n]   metres (1e30 when the query has no pair at all)
 *   min_pair[n]       pair index (cc_query_get_pairs) attaining it, -1 when none; may be NULL
 *   closest_points    [n][6] world coordinates: the point on the query link, the point on the other link; may be NULL
 * With both tolerances 0 the value is the exact minimum evaluated in FP32 (a few 1e-6 m). */
int cc_distance_batch(cc_query* q, const double* dof_values, int n, int flags, double max_distance, double rel_err,
                      double abs_err, double* min_distance, int32_t* min_pair, double* closest_points);
/* Same for given link transforms [n][nlinks][8] (single-state queries: CheckCollision(body, report) with CO_Distance). */
int cc_distance_poses(cc_query* q, const double* link_transforms8, int n, int flags, double max_distance, double rel_err,
                      double abs_err, double* min_distance, int32_t* min_pair, double* closest_points);

/* ---- CO_UseTolerance (collisionPQP.h:650-657: PQP_Tolerance per visited pair, report->numWithinTol).  within_mask
 * [n][mask_words]: bit p set <=> pair p (cc_query_get_pairs) is no farther apart than `tolerance` metres (distance <=
 * tolerance, pqp/PQP.cpp:1056; colliding pairs included).  The plugin turns the bits into numWithinTol the way
 * CheckCollisionP does: the number of other bodies with at least one pair within tolerance (collisionPQP.h:443-497). */
int cc_tolerance_batch(cc_query* q, const double* dof_values, int n, int flags, double tolerance, uint32_t* within_mask);
int cc_tolerance_poses(cc_query* q, const double* link_transforms8, int n, int flags, double tolerance, uint32_t* within_mask);

/* ---- CO_Contacts (collisionPQP.h:578-603: PQP_ALL_CONTACTS per link pair, one CONTACT per intersecting triangle pair through
 * OpenRAVE::geometry::TriTriCollision, geometry.h:1817-1912) for ONE set of link transforms [nlinks][8].  Up to max_contacts
 * records, sorted by (pair, triangle a, triangle b):
 *   out_pair[i]           pair index (cc_query_get_pairs)
 *   out_triangles[i][2]   triangle index inside the query link's mesh / the other link's mesh (the ids PQP reports)
 *   out_has_contact[i]    TriTriCollision's return value for that triangle pair (0: it yields no CONTACT, as in the reference)
 *   out_pos_norm[i][6]    contact position and normal (world), valid where out_has_contact[i]; depth is 0 (collisionPQP.h:602)
 * *out_count = number of intersecting triangle pairs found (if it exceeds max_contacts the list was cut).  The last two
 * outputs may be NULL.  cc_tri_tri_contact is the contact routine itself (3 x xyz per triangle), returns 0 / 1. */
int cc_contacts_poses(cc_query* q, const double* link_transforms8, int flags, int max_contacts, int* out_count, int32_t* out_pair,
                      int32_t* out_triangles, uint8_t* out_has_contact, double* out_pos_norm);
int cc_tri_tri_contact(const double* u9, const double* v9, double* pos3, double* norm3);

/* Page-locked host buffers for the host-buffer entry points (cc_check_batch, cc_check_poses, cc_fk_batch).  Those calls
 * cut a big batch into chunks and copy each chunk's inputs / results on the chunk's own stream; with buffers from
 * cc_host_alloc the copies overlap the kernels of the neighbouring chunk.  Ordinary (pageable) memory stays valid
 * everywhere, it only serialises the copies.  The plugin keeps one set per checker for the planners' edge batches
 * (the std::vector<dReal> the per-sample loop fills, planningutils.cpp:3405-3454). */
int cc_host_alloc(size_t bytes, void** out);
int cc_host_free(void* p);

/* Wait for everything enqueued by cc_check_batch_device on this scene's device and report asynchronous
 * failures (CC_ERR_NOMEM if a traversal overflowed its spill area). */
int cc_scene_synchronize(cc_scene* s);
/* Conservative margin of the FP32 triangle test in metres: two triangles closer than this on every
 * separating-axis candidate count as touching.  0 (default) = the reference's strict test (pqp/PQP.cpp:366-367). */
int cc_scene_set_tri_epsilon(cc_scene* s, double eps);

/* Certified free space (on by default).  Once a bound query has seen 16k batched configurations the library looks at the
 * link pairs whose relative pose depends on only one or two active DOFs (everything the torso carries against the
 * environment, a shoulder next to a wall, an upper arm next to a forearm): the exact distance on a grid over those DOFs
 * (cc_distance_kernel) minus the largest displacement any point can make inside a grid cell certifies cells
 * collision-free.  A pair certified in every cell can never collide inside the joint limits and leaves the batched pair
 * list (hoisting, generalised from pose-invariant pairs); a partly certified pair keeps a bit table that the batched entry
 * consults after the broadphase instead of descending the trees.  Verdicts and reported pairs are unchanged (a skipped pair
 * is at least 3e-4 m from contact).  Certificates are proven on the snapshot the query captured (cc_query_create /
 * cc_query_refresh): if any body was moved since (cc_body_set_link_transforms, cc_body_set_dof_values) the build waits for the
 * caller's cc_query_refresh.  cc_query_get_table_count reports the pairs with a table, the pairs removed and the
 * wall time the last build took (any pointer may be NULL). */
int cc_scene_set_pair_tables(cc_scene* s, int enabled);
int cc_query_get_table_count(const cc_query* q, int* out_tables, int* out_dropped, double* out_build_ms);
/* What the last certificate build proved, one record per pair that got a grid certificate (the reach certificates, which need
 * no grid, are not listed).  Cell (i0, i1) of the grid covers q[col0] in [lo0 + i0 width0, lo0 + (i0+1) width0] (and the same
 * for col1); it is certified when the exact distance at its centre exceeds delta + margin, delta = (k0 width0 + k1 width1) / 2
 * being the farthest any point of one link can move relative to the other inside the cell.  min_slack = the smallest
 * distance - delta - margin over the certified cells, attained in cell `tightest_cell` (= i0 n1 + i1): where a soundness
 * test should aim (tests/test_pair_tables.py). */
typedef struct cc_certificate {
    int32_t pair;                  /* pair index (cc_query_get_pairs) */
    int32_t col0, col1;            /* columns of q; col1 = -1: one DOF */
    int32_t n0, n1;                /* grid size */
    int32_t certified_cells;       /* of n0 n1 */
    int32_t tightest_cell;
    int32_t dropped;               /* 1: certified in every cell, the pair left the batched pair list; 0: it keeps a bit table */
    double lo0, width0, lo1, width1;
    double k0, k1;                 /* metres per unit of q[col0], q[col1] */
    double delta, margin;
    double min_slack, min_distance;
} cc_certificate;
int cc_query_get_certificates(const cc_query* q, cc_certificate* out, int max_records, int* out_count);

/* Diagnostics: record, for every configuration of the following batches on this query, how many traversal
 * iterations (32 box-pair tests each) the device spent on it; read the last batch's trace back with
 * cc_query_get_work_trace (synchronises).  Used by the load-balance studies in profiles/. */
int cc_query_trace_work(cc_query* q, int enabled);
int cc_query_get_work_trace(cc_query* q, uint32_t* out_iterations, int n);

/* Diagnostics: the per-chunk timeline of the NEXT batched calls.  A big batch is cut into chunks that alternate between two
 * internal streams ("lanes") so that the tail of one chunk's persistent traversal kernel overlaps the next chunk's kernels;
 * with tracing on, three timed events per chunk are recorded on the chunk's lane.  cc_scene_get_timeline (synchronises)
 * returns, for every chunk of the last traced call, milliseconds since the call's start event: {the lane became free for this
 * chunk, its FK kernel ended, its traversal kernel ended} and the chunk's lane.  Overlapping [FK end, traversal end] intervals
 * of consecutive chunks are the multi-stream overlap bench.py's roofline accounting relies on (profiles/ holds the records). */
int cc_scene_trace_timeline(cc_scene* s, int enabled);
int cc_scene_get_timeline(cc_scene* s, double* out_ms /* [chunks][3] */, int32_t* out_lane /* [chunks], may be NULL */,
                          int max_chunks, int* out_chunks);

/* Measured FP32 FMA throughput of a device in TFLOP/s (a short FMA-chain kernel): the denominator of the compute-side
 * roofline figure bench.py reports next to the memory-side one. */
int cc_device_fp32_peak(int cuda_device, double* out_tflops);

/* ---- statistics (PQP_CollideResult::NumBVTests/NumTriTests analogue, pqp/PQP.cpp:489,500) */
int cc_scene_set_counters(cc_scene* s, int enabled);
int cc_scene_get_stats(const cc_scene* s, cc_stats* out);


/* Diagnostic (no device needed): the bounding-volume hierarchy cc_scene_commit builds for one triangle mesh — what replaces
 * PQP_Model::EndModel / build_model (pqp/PQP.cpp:169-241, pqp/Build.cpp:344-551) on the device side.  vertices: [nvertices][3],
 * triangles: [ntriangles][3] vertex indices.  out_nodes (may be NULL): 8 x 32-bit words per node {cx, cy, cz, left (int32),
 * ex, ey, ez, size}: link-frame AABB centre and half extents; left >= 0: index of the first child (the second is left + 1),
 * left < 0: leaf holding triangle ~left; node 0 is the root; *out_count = 2 ntriangles - 1. */
int cc_debug_build_bvh(const double* vertices, int nvertices, const int32_t* triangles, int ntriangles, float* out_nodes,
                       int max_nodes, int* out_count);

/* Diagnostic (no device needed): the box test of the broadphase kernel on the host, compiled from the same source as the kernel.
 * boxes_a / boxes_b: [n][8] floats {cx cy cz . ex ey ez .} (the FK kernel's world AABBs, base-centred).  out[i]: bit 0 = the
 * boxes overlap (|ca - cb| <= ea + eb on all axes, evaluated in double), bit 1 = the kernel's 15-bit fixed-point test says
 * so.  The kernel is conservative when bit 0 implies bit 1 (the reference has no broadphase: collisionPQP.h:432-498 visits
 * every link pair, so a pair may only ever be dropped when its boxes are disjoint). */
int cc_debug_broad_overlap(const float* boxes_a, const float* boxes_b, int n, uint8_t* out);

/* Diagnostic (no device needed): what a bound query treats as fixed for the length of a batch — the host rule behind
 * cc_query_info.npairs_hoisted.  joints / affects as for cc_body_set_kinematics (affects: [ndof][nlinks]); active: the active DOF
 * indices; base_moves: affine DOFs are active.  out_moves[l] = 1 when link l's world pose depends on the rows of q (an active DOF
 * affects it, KinBody::DoesAffect kinbody.cpp:4787-4818; the base moves; a CC_MIMIC_EXTERNAL joint sits above it);
 * out_invariant[p] = 1 when the pose of link pairs[2p+1] in the frame of link pairs[2p] cannot change: no joint on the tree path
 * between the two is driven by an active DOF, directly, through a mimic equation, or by the caller.  An environment pair is
 * evaluated once at bind time when its link does not move, a self pair when it is invariant; the reference evaluates every pair
 * at every configuration (collisionPQP.h:364-417), so either flag may only be set when the verdict cannot change. */
int cc_debug_motion_classes(const cc_joint* joints, int njoints, int nlinks, int ndof, const uint8_t* affects, const int32_t* active,
                            int nactive, int base_moves, const int32_t* pairs, int npairs, uint8_t* out_moves, uint8_t* out_invariant);

/* Diagnostic (no device needed): the FK chain program a query compiles for the device — joints in depth-first order, everything
 * rigidly attached to a varying joint's child folded into one constant per link, pose-table slots and parked poses assigned — run
 * on the host in double precision with the device's operation order.  Arguments as for cc_body_set_kinematics /
 * cc_body_set_dof_values / cc_query_create; base_pose = link 0's pose {quat wxyz, xyz, 0}; wanted_links[nlinks] (NULL: all) marks
 * the links whose poses are needed (the compiler drops chains nobody needs); walk 0 = the order of cc_fk_kernel (running
 * transform + parked poses), 1 = cc_fk2_kernel (pose table).  q: [n][nactive + external columns]; out_poses: [n][nlinks][8],
 * zeros for links that are not wanted; out_counts (may be NULL): {steps, varying steps, pose-table slots, parked poses}.
 * Restates KinBody::SetDOFValues' joint loop (kinbody.cpp:2286-2481): tests/test_fk_program.py against the harness. */
int cc_debug_fk_program(const cc_joint* joints, int njoints, int nlinks, int ndof, const double* dofvals, const double* base_pose,
                        const int32_t* active, int nactive, const uint8_t* wanted_links, int walk, const double* q, int64_t n,
                        double* out_poses, int32_t* out_counts);

/* Diagnostics (no device needed) of the mimic-equation evaluator.  cc_debug_mimic_eval: out[i] = equation(values[i][0..nnames)),
 * err[i] = fparser's EvalError code (0 fine, 1 division by zero, 2 sqrt, 3 log, 4 trig domain) or 5 for a power / a result that
 * is NaN or infinite, which fparser reports as one of its codes or not at all depending on how it rewrote the power; out[i] is then 0.
 * cc_debug_expand_rows: cc_query_expand_rows on explicit arguments — equations[njoints] holds the text for CC_MIMIC_EXTERNAL joints
 * and NULL elsewhere, dofvals the body's current DOF values, active the active DOF indices. */
int cc_debug_mimic_eval(const char* equation, const char* const* names, int nnames, const double* values, int64_t n, double* out, int32_t* err);
int cc_debug_expand_rows(const cc_joint* joints, int njoints, int ndof, const double* dofvals, const char* const* equations,
                         const char* const* joint_names, const int32_t* active, int nactive, int naffine, const double* q_in, int64_t n,
                         double* q_out);

/* ==== planners on the batched entry (BASELINE.json configs[2]) =========================================
 * The reference's RRT family spends its time in PlannerParameters::CheckPathAllConstraints (planner.h:373-379) ->
 * DynamicsCollisionConstraint::Check (planningutils.cpp:2586-3629), one sample at a time.  These entry points run the same
 * planners with every edge check going through cc_check_batch:
 *   - one SpatialTree::Extend (plugins/rplanners/rplanners.h:308-460) is a ray of up to 100 steps whose configurations do
 *     not depend on any collision result, so the whole ray is checked in one batch and the first failing step decides how
 *     many nodes are added (identical outcome to the serial loop);
 *   - `nproblems` independent planning problems advance in lockstep: each round gathers the pending ray / shortcut of every
 *     live problem into ONE cc_check_batch call.
 * Replaces, per problem: BirrtPlanner::InitPlan/PlanPath/_ExtractPath (plugins/rplanners/rrt.h:279-659),
 * RrtPlanner::_SimpleOptimizePath (rrt.h:151-240) and the default post-processor LinearSmoother::_OptimizePath /
 * _OptimizePathSingleDOF / SegmentFeasible (plugins/rplanners/linearsmoother.cpp:279-559, 1235-1263; retiming is not part
 * of this path). */
typedef struct cc_planner cc_planner;

/* The configuration space PlannerParameters::SetRobotActiveJoints builds (planner.cpp:650-697): limits, resolutions,
 * DOF weights of SimpleDistanceMetric (planningutils.cpp:4705-4723) and the circular mask of SubtractDOFValues
 * (kinbody.cpp:1357-1394).  All arrays have ndof entries and are copied. */
typedef struct cc_cspace {
    int32_t ndof;
    int32_t reserved;
    const double* lower;
    const double* upper;
    const double* resolution;
    const double* weights;
    const uint8_t* circular;
} cc_cspace;

typedef struct cc_birrt_params {
    double step_length;              /* PlannerParameters::_fStepLength, default 0.04 (planner.cpp:315) */
    double goal_bias;                /* BirrtPlanner::_fGoalBiasProb = 0.01 (rrt.h:315) */
    double point_tolerance;          /* TrajectoryTimingParameters::_pointtolerance = 0.2 (plannerparameters.h:455) */
    int32_t max_iterations;          /* _nMaxIterations, <= 0 -> 10000 (rrt.h:358-359) */
    int32_t path_optimize_iterations;/* _ExtractPath calls _SimpleOptimizePath(path, 10) (rrt.h:647) */
    int32_t smoother_iterations;     /* linearsmoother _nMaxIterations, reference default 100; 0 = no post-processing */
    int32_t max_extend_steps;        /* Extend's "avoid infinite loops" bound, 100 (rplanners.h:324) */
    int32_t speculate_steps;         /* steps of a ray sent per round: 0 the whole ray, n > 0 fixed chunks, n < 0 adaptive (|n| first,
                                      * doubling while the chunks pass); the plans are the same for every value */
    int32_t post_processor;          /* cc_post_processor: PlannerParameters::_sPostProcessingPlanner */
    int32_t minimum_goal_paths;      /* RRTParameters::_minimumgoalpaths (plannerparameters.h:740): goals to connect to before the
                                      * shortest path is returned; <= 1 -> 1 (the default).  BiRRT only (rrt.h:495-531). */
    int32_t max_planning_time_ms;    /* PlannerParameters::_nMaxPlanningTime (planner.h; rrt.h:414-420): stop the search this long after
                                      * the call started; <= 0: no limit.  BiRRT only. */
} cc_birrt_params;

/* The path optimiser that runs on the raw plan (`_sPostProcessingPlanner`): the default "linearsmoother"
 * (planner.cpp:28, 315; plugins/rplanners/linearsmoother.cpp) or "shortcut_linear" (ShortcutLinearPlanner,
 * plugins/rplanners/linearshortcutadvanced.cpp:76-569: _SubsampleTrajectory, _OptimizePath, _OptimizePathOneDOF — the one
 * rmanipulation/basemanipulation.cpp:788 selects).  Both run `smoother_iterations` iterations per pass. */
enum cc_post_processor { CC_POST_LINEARSMOOTHER = 0, CC_POST_SHORTCUT_LINEAR = 1,
                         CC_POST_LINEARSMOOTHER_SHORTCUTS = 2 /* "linearsmoother 0": _nUseSingleDOFSmoothing = 0, shortcuts only (linearsmoother.cpp:25-26, 241-244) */ };

enum cc_plan_status { CC_PLAN_FAILED = 0, CC_PLAN_SOLVED = 1, CC_PLAN_BAD_INITIAL = 2, CC_PLAN_BAD_GOAL = 3,
                      CC_PLAN_INTERRUPTED = 4 /* PS_Interrupted: a progress callback answered CC_PA_INTERRUPT */ };

typedef struct cc_plan_result {
    int32_t status;                  /* cc_plan_status */
    int32_t iterations;              /* main-loop passes (the reference logs iter/3) */
    int32_t nodes_forward, nodes_backward;
    int32_t waypoints_raw;           /* after _ExtractPath (incl. _SimpleOptimizePath) */
    int32_t waypoints;               /* after the linear smoother */
    int32_t rounds;                  /* lockstep rounds this problem took part in */
    int32_t shortcuts;               /* successful shortcuts of the shortcut_linear post-processor (its return values) */
    double length_raw, length;       /* sum of SimpleDistanceMetric over consecutive waypoints */
    uint64_t configs_submitted;      /* configurations sent to the checker for this problem */
    uint64_t configs_serial;         /* configurations the serial per-sample loop would have checked (up to each first failure) */
    int32_t start_index, goal_index; /* GOALPATH::startindex / goalindex (rrt.h:296-302): which initial configuration (rank among the
                                      * valid ones) and which goal (index as given) the plan connects; -1 when there is no plan */
    int32_t goal_paths;              /* goals a path was found to (_vgoalpaths.size()); the plan is the shortest of them */
    int32_t reserved;
} cc_plan_result;

/* A checker for groups of ordered segments.  Configurations q[nconfigs][ndof]; segment s owns configurations
 * [segment_begin[s], segment_begin[s+1]) in the order the reference checks them; group g owns segments
 * [group_begin[g], group_begin[g+1]).  For every group the callee reports the first segment holding a colliding
 * configuration (-1 if none), the cc_verdict_bits of that configuration, and the global index of that configuration.
 * It may stop inside a group at its first failure (what the serial reference does). */
typedef int (*cc_group_check_fn)(void* user, const double* q, int nconfigs, int ndof,
                                 const int32_t* group_begin, int ngroups, const int32_t* segment_begin, int nsegments,
                                 int32_t* first_fail_segment, int32_t* fail_verdict, int32_t* first_fail_config);

int cc_planner_create(const cc_cspace* space, cc_planner** out);
int cc_planner_destroy(cc_planner* p);
/* Edge checks through cc_check_batch on `q` (its active DOFs must be the planner's DOFs, in order). */
int cc_planner_set_query(cc_planner* p, cc_query* q);
/* Edge checks through a caller-supplied checker (the test harness plugs the CPU oracle in here). */
int cc_planner_set_checker(cc_planner* p, cc_group_check_fn fn, void* user);

/* PlannerParameters::_samplegoalfn / _sampleinitialfn (planner.h:298-311; BirrtPlanner::PlanPath, rrt.h:422-437): called once per
 * main-loop iteration of every BiRRT problem; a callback that returns 1 has written a new goal (initial) configuration to
 * q_out[ndof], which becomes a new root of the backward (forward) tree as it is — the reference does not check it either; its
 * goal index is the next free one.  With a goal sampler a problem may start without any valid goal.  NULL switches a sampler
 * off.  `problem` is the problem's position in the batch. */
typedef int (*cc_config_sampler_fn)(void* user, int problem, int ndof, double* q_out);

/* PlannerBase::RegisterPlanCallback / _CallCallbacks (planner.h:76-79, 700-716; BirrtPlanner::PlanPath, rrt.h:401-412): called at the
 * top of every main-loop iteration of every BiRRT problem with PlannerProgress::_iteration; the answer is a PlannerAction. */
enum cc_planner_action { CC_PA_NONE = 0, CC_PA_INTERRUPT = 1, CC_PA_RETURN_WITH_ANY_SOLUTION = 2 };
typedef int (*cc_progress_fn)(void* user, int problem, int iteration);
int cc_planner_set_progress_callback(cc_planner* planner, cc_progress_fn fn, void* user);
int cc_planner_set_samplers(cc_planner* planner, cc_config_sampler_fn goal_fn, cc_config_sampler_fn initial_fn, void* user);
/* The sample sequence DynamicsCollisionConstraint::Check visits for the edge q0 -> q1 (planningutils.cpp:2692-2835,
 * 3405-3454), in checking order.  interval: 0 IT_Open, 1 IT_OpenStart, 2 IT_OpenEnd, 3 IT_Closed (planner.h).  Writes at
 * most max_configs rows to out_configs (may be NULL) and returns the count in *out_count; -1 there = the edge leaves the
 * joint limits (state-setting error). */
int cc_planner_edge_samples(cc_planner* p, const double* q0, const double* q1, int interval,
                            double* out_configs, int max_configs, int* out_count);
/* ConstraintFilterReturn::_configurationtimes of the same samples (planningutils.cpp:2801-2808, 3410-3413): the interpolation
 * coefficient in [0, 1] of every sample of cc_planner_edge_samples, in the same (checking) order. */
int cc_planner_edge_sample_times(cc_planner* p, const double* q0, const double* q1, int interval, double* out_times, int max_configs,
                                 int* out_count);
/* The same for a parabolic ramp: DynamicsCollisionConstraint::Check's quadratic branch (planningutils.cpp:2660-2681, 2726-2766,
 * 2844-3319), taken when the caller passes end-point velocities and timeelapsed > 0 — the parabolic smoothers' feasibility
 * checks (plugins/rplanners/parabolicsmoother.cpp, CheckPathAllConstraints(a, b, da, db, timeelapsed, ...)).  The sampled
 * positions follow q0 + dq0 t + (dq1 - dq0) t^2 / (2 timeelapsed); out_times[i] is the ramp time of sample i (NaN for the
 * linear sub-steps the reference inserts where a step still exceeds 1.01 x a resolution) and out_kind[i] is its
 * cc_ramp_sample_kind: only CC_RAMP_REPORTED samples appear in ConstraintFilterReturn::_configurations.  All samples are in
 * checking order; the tail sub-steps (between the last ramp sample and q1) are checked BEFORE that last ramp sample (:3207-3296).  *out_count = -1: CFO_StateSettingError (no root, the ramp needs steps 100x finer than the resolutions, ...). */
int cc_planner_ramp_samples(cc_planner* p, const double* q0, const double* q1, const double* dq0, const double* dq1,
                            double timeelapsed, int interval, double* out_configs, double* out_times, uint8_t* out_kind,
                            int max_configs, int* out_count);
enum cc_ramp_sample_kind { CC_RAMP_SUBSTEP = 0, CC_RAMP_REPORTED = 1, CC_RAMP_TAIL_SUBSTEP = 2, CC_RAMP_END_UNREPORTED = 3 };
/* PlannerParameters::CheckPathAllConstraints (planner.h:373-379) for `nedges` edges in ONE checker call — the entry a caller of
 * DynamicsCollisionConstraint::Check (planningutils.cpp:2586-3629) switches to.  q0, q1: [nedges][ndof].  dq0, dq1 ([nedges][ndof])
 * and timeelapsed ([nedges]) may be NULL: positions-only linear edges; an edge with timeelapsed > 0 is sampled as a parabolic
 * ramp (cc_planner_ramp_samples).  options: CFO_CheckEnvCollisions (1) | CFO_CheckSelfCollisions (2) are required,
 * CFO_CheckWithPerturbation (0x10000) also checks every sample at +/- perturbation x resolution (planningutils.cpp:2344-2366,
 * default perturbation 0.1).  Per edge: the reference's return code (0, 1 environment, 2 self, 0x80000000 state setting error),
 * the index of the failing sample in checking order, ConstraintFilterReturn::_fTimeWhenInvalid, and (invalid_values, may be
 * NULL, [nedges][ndof]) _invalidvalues. */
typedef struct cc_edge_result {
    int32_t return_code;
    int32_t first_invalid;       /* -1: every sample passed */
    int32_t nsamples;            /* samples the edge was discretised into (0 after a state setting error) */
    int32_t perturbed;           /* which state of the failing sample collided: 0 itself, 1 +perturbation, 2 -perturbation */
    double time_when_invalid;
} cc_edge_result;
int cc_planner_check_edges(cc_planner* p, int nedges, const double* q0, const double* q1, const double* dq0, const double* dq1,
                           const double* timeelapsed, int interval, int options, double perturbation, cc_edge_result* results,
                           double* invalid_values);
/* BiRRT + path shortcutting for nproblems (start, goal) pairs in lockstep.  starts/goals: [nproblems][ndof];
 * seeds[nproblems] = PlannerParameters::_nRandomGeneratorSeed per problem.  results[nproblems] is filled; the paths stay in
 * the planner until the next call and are read with cc_planner_get_path. */
int cc_birrt_plan_batch(cc_planner* p, const cc_birrt_params* params, int nproblems, const double* starts,
                        const double* goals, const uint32_t* seeds, cc_plan_result* results);
/* BiRRT with several initial and several goal configurations per problem (RRTParameters::vinitialconfig / vgoalconfig holding
 * more than one configuration, rrt.h:72-88, 327-343 — e.g. one goal per IK solution): problem k owns the initial configurations
 * [start_begin[k], start_begin[k+1]) of `starts` and the goals [goal_begin[k], goal_begin[k+1]) of `goals`.  Configurations that
 * fail the constraints are skipped (a goal keeps its index); the goal to extend toward is drawn like rrt.h:443-466; the search
 * ends at the first connection (_minimumgoalpaths = 1) and reports which start / goal it joined. */
int cc_birrt_plan_batch_multi(cc_planner* p, const cc_birrt_params* params, int nproblems, const int32_t* start_begin,
                              const double* starts, const int32_t* goal_begin, const double* goals, const uint32_t* seeds,
                              cc_plan_result* results);
/* The same for BasicRrtPlanner (plugins/rplanners/rrt.h:693-938; RaveCreatePlanner(env, "basicrrt")): one forward tree, the goal
 * sampled at the first iteration and then with probability goal_bias (0.05 in the reference), SpatialTree::Extend rays checked
 * in one batch each, success when a ray connects within 2 x step_length of the goal; extent_type = BasicRRTParameters::
 * _nRRTExtentType (1: a ray stops at its first new node).  The path ends at the connecting node like the reference's. */
int cc_basicrrt_plan_batch(cc_planner* p, const cc_birrt_params* params, int extent_type, int nproblems, const double* starts,
                           const double* goals, const uint32_t* seeds, cc_plan_result* results);
/* which: 0 = smoothed path, 1 = raw path.  Copies up to max_waypoints rows; *out_waypoints = the path's length. */
int cc_planner_get_path(const cc_planner* p, int problem, int which, double* out, int max_waypoints, int* out_waypoints);
/* Totals of the last cc_birrt_plan_batch: rounds, checker calls, configurations submitted, serial-equivalent configurations. */
int cc_planner_get_totals(const cc_planner* p, uint64_t* out4);

/* Diagnostic (no device needed): the per-DOF displacement rates K the free-space certificates rest on, for the pair (link_a,
 * link_b) of one joint tree (link_b = -1: the other link is static in the world).  K_d bounds, for EVERY value of every joint,
 * how far any point of one link moves in the other's frame per unit of DOF d.  *out_n = number of varying active DOFs between
 * the two links (1 or 2; out_dofs / out_k filled), 0 when such a pair gets no certificate. */
int cc_debug_pair_rates(const cc_joint* joints, int njoints, int nlinks, const int32_t* col_of_dof, int ndof, int link_a, int link_b,
                        double radius_a, double radius_b, int32_t* out_dofs, double* out_k, int* out_n);

#ifdef __cplusplus
}
#endif
#endif /* CUDACOLLISION_H */
