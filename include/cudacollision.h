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
    int32_t mimic_dof;     /* DOF the mimic equation reads, -1 if not mimic */
    int32_t is_static;     /* Joint::IsStatic() */
    int32_t is_circular;   /* Joint::IsCircular(0) */
    int32_t reserved;
    double mimic_a, mimic_b;
    double lower, upper;   /* Joint::GetLimits */
    double left[8];
    double right[8];
    double axis[4];        /* xyz + pad */
} cc_joint;

/* Flags of the batched entry. */
enum cc_batch_flags {
    CC_BATCH_ENV = 1,        /* Environment::CheckCollision(robot) */
    CC_BATCH_SELF = 2,       /* KinBody::CheckSelfCollision(robot) */
    CC_BATCH_ALL_PAIRS = 4   /* no early exit: fill the complete per-pair bitmask (report != NULL path) */
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
int cc_scene_destroy(cc_scene* s);

/* ---- InitKinBody / RemoveKinBody (collisionchecker.h:233-237; pqprave: collisionPQP.h:89-122) */
int cc_scene_add_body(cc_scene* s, const char* name, int nlinks, int is_robot, int* out_body);
int cc_scene_remove_body(cc_scene* s, int body);
/* Link::GetCollisionData(): link-frame vertices (xyz doubles) + triangle indices (TriMesh, openrave.h:2501-2505) */
int cc_body_set_link_mesh(cc_scene* s, int body, int link, const char* link_name,
                          const double* vertices_xyz, int nvertices, const int32_t* triangles, int ntriangles);
/* Link::IsEnabled() for every link */
int cc_body_set_link_enabled(cc_scene* s, int body, const uint8_t* enabled);
/* KinBody::GetLinkTransformations(): [nlinks][8] */
int cc_body_set_link_transforms(cc_scene* s, int body, const double* transforms8);
/* KinBody::IsAttached (grabbed / attached bodies are never checked against each other) */
int cc_body_set_attached(cc_scene* s, int body_a, int body_b, int attached);

/* ---- kinematics of a body, read off a computed KinBody (used by the batched entry only) */
int cc_body_set_kinematics(cc_scene* s, int body, int ndof, int njoints, const cc_joint* joints_in_fk_order,
                           const uint8_t* dof_affects_link /* [ndof][nlinks], KinBody::DoesAffect != 0 */);
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
int cc_query_destroy(cc_query* q);
int cc_query_get_info(const cc_query* q, cc_query_info* info);
/* pair index -> {query link, other body, other link, 0=env|1=self}, [npairs][4] */
int cc_query_get_pairs(const cc_query* q, int32_t* out_quads);

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
 * `cuda_stream` (a cudaStream_t, NULL = legacy default stream) and NOT synchronised. */
int cc_check_batch_device(cc_query* q, const double* d_dof_values, int n, int flags,
                          uint32_t* d_verdict, int32_t* d_first_env, int32_t* d_first_self,
                          uint32_t* d_pair_mask, void* cuda_stream);
/* Single-state flavour: link transforms given instead of DOF values, [n][nlinks][8] doubles
 * (CheckCollision overloads read the bodies' CURRENT transforms and never run FK, collisionPQP.h:438-447). */
int cc_check_poses(cc_query* q, const double* link_transforms8, int n, int flags,
                   uint32_t* verdict, int32_t* first_env, int32_t* first_self, uint32_t* pair_mask);
/* Batched KinBody::SetDOFValues: out [n][nlinks][8] doubles (validation of the 1e-6 m / 1e-6 rad bound). */
int cc_fk_batch(cc_query* q, const double* dof_values, int n, double* out_transforms8);

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

/* Diagnostics: record, for every configuration of the following batches on this query, how many traversal
 * iterations (32 box-pair tests each) the device spent on it; read the last batch's trace back with
 * cc_query_get_work_trace (synchronises).  Used by the load-balance studies in profiles/. */
int cc_query_trace_work(cc_query* q, int enabled);
int cc_query_get_work_trace(cc_query* q, uint32_t* out_iterations, int n);

/* ---- statistics (PQP_CollideResult::NumBVTests/NumTriTests analogue, pqp/PQP.cpp:489,500) */
int cc_scene_set_counters(cc_scene* s, int enabled);
int cc_scene_get_stats(const cc_scene* s, cc_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* CUDACOLLISION_H */
