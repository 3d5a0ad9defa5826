// cc_contacts.inl — CO_Contacts on the report path (host code, included by cudacollision.cu).
//
// pqprave collides a link pair with PQP_ALL_CONTACTS and, when a report is given, turns every intersecting triangle pair into
// a CONTACT through OpenRAVE::geometry::TriTriCollision (plugins/pqprave/collisionPQP.h:578-603, include/openrave/geometry.h:
// 1817-1912).  Here the traversal kernel's all-contacts flavour enumerates the intersecting triangle pairs of one
// configuration; the contact point and normal of each are evaluated on the host in double, from the original (double)
// vertices and link poses, by the routine below.
namespace {

struct V3d { double x, y, z; };
inline V3d operator-(V3d a, V3d b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3d operator+(V3d a, V3d b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3d operator*(V3d a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot3(V3d a, V3d b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3d cross3(V3d a, V3d b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// The contact of triangle U against triangle V as geometry.h's TriTriCollision defines it:
//   1. U must straddle V's plane (its vertices are classified by the sign of their plane distance; all on one side = none);
//   2. the two edges of U that leave its lone vertex are cut by V's plane: that is the segment in which U crosses the plane;
//   3. the segment is clipped against the three edge half-planes of V (an end on the outside is moved onto the edge line,
//      both ends outside one edge = no contact);
//   4. contact position = midpoint of what is left, contact normal = V's unit normal.
bool tri_tri_contact(const V3d u[3], const V3d v[3], V3d& pos, V3d& norm)
{
    const V3d vedge[3] = {v[1] - v[0], v[2] - v[1], v[0] - v[2]};
    const V3d vnorm = cross3(vedge[2], vedge[0]);
    const double vw = -dot3(vnorm, v[0]);
    int side = 0;
    for (int i = 0; i < 3; ++i) if (dot3(vnorm, u[i]) + vw > 0) side |= 1 << i;
    if (side == 0 || side == 7) return false;
    // the vertex alone on its side of the plane
    const int lone = (side == 1 || side == 6) ? 0 : ((side == 2 || side == 5) ? 1 : 2);
    const V3d pu = u[lone];
    const V3d e1 = u[lone == 0 ? 1 : 0] - pu;
    const V3d e2 = u[lone == 2 ? 1 : 2] - pu;
    const double t = dot3(vnorm, pu) + vw;
    V3d p1 = pu - e1 * (t / dot3(vnorm, e1));
    V3d p2 = pu - e2 * (t / dot3(vnorm, e2));
    for (int i = 0; i < 3; ++i) {
        const V3d q1 = p1 - v[i], q2 = p2 - v[i];
        const V3d out = cross3(vedge[i], vnorm);          // points out of V across edge i
        const double t1 = dot3(q1, out), t2 = dot3(q2, out);
        if (t1 >= 0 && t2 >= 0) return false;
        if (t1 > 0 && t2 < 0) { const V3d dq = q2 - q1; p1 = p1 - dq * (t1 / dot3(dq, out)); }
        else if (t1 < 0 && t2 > 0) { const V3d dq = q1 - q2; p2 = p2 - dq * (t2 / dot3(dq, out)); }
    }
    pos = (p1 + p2) * 0.5;
    const double len = sqrt(dot3(vnorm, vnorm));
    norm = vnorm * (1.0 / len);
    return true;
}

// RaveTransform * point (geometry.h:455-464: rotate by the quaternion, add the translation)
V3d xf_apply(const double* x, const double* p)
{
    const double q0 = x[0], q1 = x[1], q2 = x[2], q3 = x[3];
    const double xx = 2 * q1 * q1, xy = 2 * q1 * q2, xz = 2 * q1 * q3, xw = 2 * q1 * q0;
    const double yy = 2 * q2 * q2, yz = 2 * q2 * q3, yw = 2 * q2 * q0, zz = 2 * q3 * q3, zw = 2 * q3 * q0;
    return {(1 - yy - zz) * p[0] + (xy - zw) * p[1] + (xz + yw) * p[2] + x[4],
            (xy + zw) * p[0] + (1 - xx - zz) * p[1] + (yz - xw) * p[2] + x[5],
            (xz - yw) * p[0] + (yz + xw) * p[1] + (1 - xx - yy) * p[2] + x[6]};
}

}  // namespace

extern "C" int cc_tri_tri_contact(const double* u9, const double* v9, double* pos3, double* norm3)
{
    if (!u9 || !v9 || !pos3 || !norm3) return -1;
    V3d u[3] = {{u9[0], u9[1], u9[2]}, {u9[3], u9[4], u9[5]}, {u9[6], u9[7], u9[8]}};
    V3d v[3] = {{v9[0], v9[1], v9[2]}, {v9[3], v9[4], v9[5]}, {v9[6], v9[7], v9[8]}};
    V3d p, n;
    if (!tri_tri_contact(u, v, p, n)) return 0;
    pos3[0] = p.x; pos3[1] = p.y; pos3[2] = p.z; norm3[0] = n.x; norm3[1] = n.y; norm3[2] = n.z;
    return 1;
}

extern "C" int cc_contacts_poses(cc_query* q, const double* T8, int flags, int max_contacts, int* out_count, int32_t* out_pair,
                                 int32_t* out_triangles, uint8_t* out_has_contact, double* out_pos_norm)
{
    if (!q || !T8 || !out_count || max_contacts < 0 || (max_contacts > 0 && (!out_pair || !out_triangles)))
        return fail(CC_ERR_INVALID, "bad arguments");
    cc_scene* S = q->scene;
    if (!S->committed) return fail(CC_ERR_STATE, "scene changed since the query was bound");
    CC_CUDA(cudaSetDevice(S->device));
    const Plan& pl = q->full;
    *out_count = 0;
    if (pl.pairs.empty()) return CC_OK;
    // ---- enumerate the intersecting triangle pairs on the device
    const int cap = std::max(max_contacts, 1);
    DevBuf<uint4> d_out;
    DevBuf<uint32_t> d_count;
    CC_CUDA(d_out.reserve(cap));
    CC_CUDA(d_count.reserve(1));
    CC_CUDA(cudaMemset(d_count.p, 0, sizeof(uint32_t)));
    CC_CUDA(q->d_q.reserve((size_t)q->nq * 8));
    CC_CUDA(cudaMemcpy(q->d_q.p, T8, (size_t)q->nq * 8 * sizeof(double), cudaMemcpyHostToDevice));
    CC_CUDA(q->d_mask.reserve(q->mask_words));
    q->contact_out = d_out.p; q->contact_count = d_count.p; q->contact_cap = cap;
    int rc = launch(q, pl, false, q->d_q.p, 1, (flags & (CC_BATCH_ENV | CC_BATCH_SELF)) | CC_BATCH_ALL_PAIRS | ccdev::kFlagPoseMode, nullptr,
                    nullptr, nullptr, q->d_mask.p, nullptr, 0);
    q->contact_out = nullptr; q->contact_count = nullptr; q->contact_cap = 0;
    if (rc != CC_OK) return rc;
    CC_CUDA(cudaStreamSynchronize(0));
    rc = check_err_flag(S);
    if (rc != CC_OK) return rc;
    S->stats.configs -= 1;
    uint32_t produced = 0;
    CC_CUDA(cudaMemcpy(&produced, d_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    const int n = (int)std::min<uint32_t>(produced, (uint32_t)max_contacts);
    std::vector<uint4> rec(std::max(n, 1));
    if (n > 0) CC_CUDA(cudaMemcpy(rec.data(), d_out.p, (size_t)n * sizeof(uint4), cudaMemcpyDeviceToHost));
    rec.resize(n);
    std::sort(rec.begin(), rec.end(), [](const uint4& a, const uint4& b) {
        if (a.y != b.y) return a.y < b.y;
        if (a.z != b.z) return a.z < b.z;
        return a.w < b.w;
    });
    *out_count = (int)produced;              // > max_contacts tells the caller the list was cut
    // ---- contact point / normal of each, in double, from the original vertices and the poses of this call
    for (int i = 0; i < n; ++i) {
        const int p = (int)rec[i].y;
        const int a = q->quads[4 * p], ob = q->quads[4 * p + 1], j = q->quads[4 * p + 2], cls = q->quads[4 * p + 3];
        const HLink& LA = q->qlink(a);
        const HLink& LB = cls == 1 ? q->qlink(j) : S->bodies[ob]->links[j];
        const int ta = (int)rec[i].z - LA.tri_base, tb = (int)rec[i].w - LB.tri_base;
        out_pair[i] = p;
        out_triangles[2 * i] = ta; out_triangles[2 * i + 1] = tb;
        if (ta < 0 || ta >= LA.ntris() || tb < 0 || tb >= LB.ntris()) return fail(CC_ERR_STATE, "contact record names a triangle outside its link");
        if (!out_pos_norm && !out_has_contact) continue;
        const double* xa = T8 + 8 * (size_t)a;
        const double* xb = cls == 1 ? T8 + 8 * (size_t)j : LB.xf;
        V3d u[3], v[3];
        for (int k = 0; k < 3; ++k) {
            u[k] = xf_apply(xa, &LA.verts[3 * (size_t)LA.tris[3 * ta + k]]);
            v[k] = xf_apply(xb, &LB.verts[3 * (size_t)LB.tris[3 * tb + k]]);
        }
        V3d pos = {0, 0, 0}, norm = {0, 0, 0};
        const bool has = tri_tri_contact(u, v, pos, norm);
        if (out_has_contact) out_has_contact[i] = has ? 1 : 0;
        if (out_pos_norm) {
            double* o = out_pos_norm + 6 * (size_t)i;
            o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = norm.x; o[4] = norm.y; o[5] = norm.z;
        }
    }
    return CC_OK;
}
