// cc_certify.inl — certified free space: generalised hoisting for pairs that depend on one or two active DOFs, and the reach
// certificate for environment pairs.  Included by cudacollision.cu (host code; uses its scene / query / plan
// structures, launch_distance and build_plan).
// ================================================================================================ certified free-space tables
// Near-miss pairs — a shoulder next to a wall, an upper arm next to a forearm — pass the AABB test in every configuration and
// cost a full descent each time just to prove separation.  When the relative pose of such a pair depends on only one or two
// active DOFs, that proof can be made once: the exact distance at the centre of every cell of a grid over those DOFs
// (cc_distance_kernel), minus the largest displacement any point of one link can make relative to the other inside the cell
// (sum over the DOFs of half cell width x K_d, K_d = 1 for a prismatic joint, the largest possible distance of a point of the
// link from the joint anchor for a revolute one, bounded by chain segment lengths so that it holds for every value of the
// other joints), certifies the whole cell collision-free.  The traversal kernel skips a certified pair after its AABB test;
// everything else runs the normal narrowphase, so verdicts are unchanged.  Which pairs get a table is decided by a short
// calibration run that counts how often each pair passes the AABB test.
namespace {

double norm3(const double* t) { return sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]); }

struct ChainInfo {
    std::vector<int> parent_joint;   // [links] joint whose child is the link, -1 for the root
};

// joints from link l up to the root (child-to-parent order)
std::vector<int> path_to_root(const HBody& B, const ChainInfo& ci, int l)
{
    std::vector<int> out;
    for (int guard = 0; l >= 0 && ci.parent_joint[l] >= 0 && guard < 1 << 16; ++guard) { out.push_back(ci.parent_joint[l]); l = B.joints[ci.parent_joint[l]].parent_link; }
    return out;
}

double joint_span(const cc_joint& j)     // largest change of |child origin - parent origin| contribution of one joint
{
    double s = norm3(j.left + 4) + norm3(j.right + 4);
    // (a mimic joint's value may sit up to g_fEpsilonEvalJointLimit = 1.8e-10 outside its limits, kinbody.cpp:2326-2386)
    if (!j.is_static && j.type == CC_JOINT_PRISMATIC) s += norm3(j.axis) * (std::max(fabs(j.lower), fabs(j.upper)) + (j.mimic_dof >= 0 ? 1e-9 : 0.0));
    return s;
}

double link_radius(const HLink& L)
{
    double r2 = 0;
    for (size_t i = 0; i + 2 < L.verts.size(); i += 3) r2 = std::max(r2, L.verts[i] * L.verts[i] + L.verts[i + 1] * L.verts[i + 1] + L.verts[i + 2] * L.verts[i + 2]);
    return sqrt(r2);
}

// K_d for every active DOF on the tree path between query links a and b (b == -1: the other link is static in the world, the
// path is root -> a).  Returns false when the pair cannot be tabulated (a mimic joint on the path, a joint with a DOF that
// is active but not 1:1, more than two varying DOFs).  `moving_radius_b`: radius of the link whose points are followed.
bool pair_rates(const HBody& B, const ChainInfo& ci, const std::vector<int>& col_of_dof, int a, int b, double radius_a, double radius_b,
                std::vector<int>& dofs, std::vector<double>& K)
{
    std::vector<int> pa = path_to_root(B, ci, a), pb;
    if (b >= 0) {
        pb = path_to_root(B, ci, b);
        while (!pa.empty() && !pb.empty() && pa.back() == pb.back()) { pa.pop_back(); pb.pop_back(); }
    }
    // varying joints
    dofs.clear(); K.clear();
    std::vector<std::pair<int, bool>> varying;      // (joint, on a's branch)
    for (int side = 0; side < 2; ++side)
        for (int j : (side == 0 ? pa : pb)) {
            const cc_joint& J = B.joints[j];
            if (J.mimic_dof == CC_MIMIC_EXTERNAL) return false;          // value supplied per configuration: not a function of one DOF we know
            if (J.mimic_dof >= 0) { if (col_of_dof[J.mimic_dof] >= 0) return false; continue; }
            if (J.is_static || J.dof_index < 0) continue;
            if (col_of_dof[J.dof_index] >= 0) varying.push_back(std::make_pair(j, side == 0));
        }
    if (varying.empty() || varying.size() > 2) return false;
    // followed link = b seen from a's frame (or a seen from the world when b is static); the swapped roles give another valid
    // bound, the smaller sum is kept by the caller through two calls
    for (auto& v : varying) {
        const cc_joint& J = B.joints[v.first];
        double k;
        if (J.type == CC_JOINT_PRISMATIC) k = norm3(J.axis);
        else {
            // distance from the joint anchor (origin of parent * left) to any point of the followed link
            double len = 0;
            if (b < 0) {
                // world-static other link: follow a; the joint is on a's branch: segments below the joint towards a
                len = norm3(J.right + 4);
                for (int j : pa) { if (j == v.first) break; len += joint_span(B.joints[j]); }
                k = len + radius_a;
            } else if (!v.second) {
                // joint on b's branch, follow b: segments below the joint towards b
                len = norm3(J.right + 4);
                for (int j : pb) { if (j == v.first) break; len += joint_span(B.joints[j]); }
                k = len + radius_b;
            } else {
                // joint on a's branch, follow b as seen from a: up from the joint to the common ancestor, then down to b
                len = norm3(J.left + 4);
                bool above = false;
                for (int j : pa) { if (above) len += joint_span(B.joints[j]); if (j == v.first) above = true; }
                for (int j : pb) len += joint_span(B.joints[j]);
                k = len + radius_b;
            }
        }
        dofs.push_back(J.dof_index);
        K.push_back(k);
    }
    return true;
}

}  // namespace

extern "C" {
static int build_plan(cc_query* Q, Plan& pl, const std::vector<int>& plist, const std::vector<uint8_t>& dynamic, bool with_fk);

// Diagnostic, no device needed: the displacement rates the certificates rest on, for one link pair of a joint tree given the way
// cc_body_set_kinematics takes it.  *out_n = number of varying DOFs (1 or 2) with out_dofs / out_k filled, 0 when the pair
// cannot be tabulated.  The CPU tests check K against finite differences of the FK restatement (tests/test_pair_tables.py).
int cc_debug_pair_rates(const cc_joint* joints, int njoints, int nlinks, const int32_t* col_of_dof, int ndof, int link_a, int link_b,
                        double radius_a, double radius_b, int32_t* out_dofs, double* out_k, int* out_n)
{
    if (!joints || njoints <= 0 || nlinks <= 0 || !col_of_dof || !out_dofs || !out_k || !out_n || link_a < 0 || link_a >= nlinks ||
        link_b >= nlinks)
        return fail(CC_ERR_INVALID, "cc_debug_pair_rates: bad arguments");
    HBody B;
    B.links.resize(nlinks);
    B.ndof = ndof;
    B.joints.assign(joints, joints + njoints);
    ChainInfo ci;
    ci.parent_joint.assign(nlinks, -1);
    for (int k = 0; k < njoints; ++k) {
        if (joints[k].child_link < 0 || joints[k].child_link >= nlinks || joints[k].parent_link < 0 || joints[k].parent_link >= nlinks)
            return fail(CC_ERR_INVALID, "cc_debug_pair_rates: joint link out of range");
        ci.parent_joint[joints[k].child_link] = k;
    }
    std::vector<int> cols(col_of_dof, col_of_dof + ndof), dofs;
    std::vector<double> K;
    *out_n = 0;
    if (!pair_rates(B, ci, cols, link_a, link_b, radius_a, radius_b, dofs, K)) return CC_OK;
    *out_n = (int)dofs.size();
    for (size_t i = 0; i < dofs.size(); ++i) { out_dofs[i] = dofs[i]; out_k[i] = K[i]; }
    return CC_OK;
}
}

// (Re)upload the tables and the per-pair table indices in the order of the batched plan and of its split variants.
static int upload_pair_tables(cc_query* Q)
{
    const Plan& pl = Q->mainl;
    int valid = 0;
    for (int32_t t : Q->table_of_orig) valid += t >= 0 ? 1 : 0;
    Q->ntables = valid;
    if (Q->h_tables.empty()) return CC_OK;
    std::vector<int32_t> t(pl.orig.size(), -1);
    for (size_t i = 0; i < t.size(); ++i) t[i] = Q->table_of_orig[pl.orig[i]];
    CC_CUDA(Q->d_tables.upload(Q->h_tables));
    CC_CUDA(Q->d_table_bits.upload(Q->h_bits));
    CC_CUDA(Q->d_pair_table.upload(t));
    for (int si = 0; si < 3; ++si) {
        std::vector<int32_t> ts(pl.split[si].orig.size(), -1);
        for (size_t i = 0; i < ts.size(); ++i) ts[i] = Q->table_of_orig[pl.split[si].orig[i]];
        CC_CUDA(Q->d_pair_table_split[si].upload(ts));
    }
    return CC_OK;
}

static int build_pair_tables(cc_query* Q)
{
    cc_scene* S = Q->scene;
    HBody* B = S->bodies[Q->body];
    Q->tables_done = true;
    if (!Q->allow_tables || !S->pair_tables || getenv("CC_TUNE_NO_TABLES") || Q->p_main.empty() || Q->ncols == 0) return CC_OK;
    const int nlb = (int)B->links.size();
    const std::vector<int>& col_of_dof = Q->col_of_dof;
    if ((int)col_of_dof.size() != B->ndof) return CC_OK;
    std::vector<const cc_joint*> joint_of_dof(B->ndof, nullptr);
    for (const cc_joint& j : B->joints) if (j.dof_index >= 0 && j.mimic_dof < 0) joint_of_dof[j.dof_index] = &j;
    for (int d = 0; d < B->ndof; ++d) if (col_of_dof[d] >= 0 && !joint_of_dof[d]) return CC_OK;     // an active DOF without its own joint
    ChainInfo ci;
    ci.parent_joint.assign(nlb, -1);
    for (size_t k = 0; k < B->joints.size(); ++k) ci.parent_joint[B->joints[k].child_link] = (int)k;
    std::vector<double> radius_cache(nlb, -1.0);
    auto rad = [&](int l) { if (radius_cache[l] < 0) radius_cache[l] = link_radius(B->links[l]); return radius_cache[l]; };

    // certificates persist in the query (cc_query_refresh drops only the ones a moved body touches); a build examines the pairs
    // that have none yet
    const size_t npairs_all = Q->quads.size() / 4;
    if (Q->cert_dropped.size() != npairs_all) { Q->cert_dropped.assign(npairs_all, 0); Q->table_of_orig.assign(npairs_all, -1); Q->h_tables.clear(); Q->h_bits.clear(); Q->certs.clear(); }
    std::vector<uint8_t>& dropped = Q->cert_dropped;
    const std::vector<uint8_t> dropped_before(dropped);
    // safety margin of every certificate: FP32 error of the distance kernel (~1e-5 m) + the 1e-4 m contact band, plus the
    // conservative margin of the triangle test when the caller has set one (pairs closer than it count as touching)
    Q->cert_tri_eps = S->tri_eps;
    const double margin = 3e-4 + (double)S->tri_eps;
    // ---- reach certificate (any number of DOFs): every point of a link stays inside a ball around the anchor of the first
    // varying joint on its path from the base (a fixed point of the world: everything above that joint is fixed), of radius
    // chain length below the anchor + the link's radius.  An environment link whose world AABB is farther from the centre than
    // that radius can never touch the link.
    {
        std::vector<double> ball(4 * (size_t)nlb, -1.0);            // per link: centre xyz, radius (radius < 0: none)
        for (int l = 0; l < nlb; ++l) {
            std::vector<int> path = path_to_root(*B, ci, l);        // child-to-parent order
            int first = -1;                                         // the varying joint closest to the root
            bool ok = true;
            for (int idx = (int)path.size() - 1; idx >= 0 && ok; --idx) {
                const cc_joint& J = B->joints[path[idx]];
                if (J.mimic_dof == CC_MIMIC_EXTERNAL) { ok = false; continue; }
                if (J.mimic_dof >= 0) { if (col_of_dof[J.mimic_dof] >= 0) { if (first < 0) ok = false; } continue; }
                if (!J.is_static && J.dof_index >= 0 && col_of_dof[J.dof_index] >= 0 && first < 0) first = idx;
            }
            if (!ok || first < 0) continue;
            const cc_joint& JF = B->joints[path[first]];
            double radius = norm3(JF.right + 4) + rad(l);
            if (JF.type == CC_JOINT_PRISMATIC) radius += norm3(JF.axis) * std::max(fabs(JF.lower), fabs(JF.upper));
            for (int idx = first - 1; idx >= 0; --idx) radius += joint_span(B->joints[path[idx]]);
            // anchor = origin of (parent link pose * left); the parent link does not move with the active DOFs
            const double* x = B->links[JF.parent_link].xf;
            const double il = 2.0 / (x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
            const double q0 = x[0], q1 = x[1], q2 = x[2], q3 = x[3];
            const double Rm[9] = {1 - il * (q2 * q2 + q3 * q3), il * (q1 * q2 - q0 * q3), il * (q1 * q3 + q0 * q2),
                                  il * (q1 * q2 + q0 * q3), 1 - il * (q1 * q1 + q3 * q3), il * (q2 * q3 - q0 * q1),
                                  il * (q1 * q3 - q0 * q2), il * (q2 * q3 + q0 * q1), 1 - il * (q1 * q1 + q2 * q2)};
            const double* tl = JF.left + 4;
            for (int r = 0; r < 3; ++r) ball[4 * l + r] = x[4 + r] + Rm[3 * r] * tl[0] + Rm[3 * r + 1] * tl[1] + Rm[3 * r + 2] * tl[2];
            ball[4 * l + 3] = radius;
        }
        for (int p : Q->p_main) {
            if (Q->quads[4 * p + 3] != 0) continue;
            const int a = Q->quads[4 * p];
            if (a >= nlb || ball[4 * a + 3] < 0) continue;
            const HLink& E = S->bodies[Q->quads[4 * p + 1]]->links[Q->quads[4 * p + 2]];
            if (E.nodes.empty()) continue;
            // world AABB of the environment link's (already inflated) root box, in double
            const double* x = E.xf;
            const double il = 2.0 / (x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
            const double q0 = x[0], q1 = x[1], q2 = x[2], q3 = x[3];
            const double Rm[9] = {1 - il * (q2 * q2 + q3 * q3), il * (q1 * q2 - q0 * q3), il * (q1 * q3 + q0 * q2),
                                  il * (q1 * q2 + q0 * q3), 1 - il * (q1 * q1 + q3 * q3), il * (q2 * q3 - q0 * q1),
                                  il * (q1 * q3 - q0 * q2), il * (q2 * q3 + q0 * q1), 1 - il * (q1 * q1 + q2 * q2)};
            double d2 = 0;
            for (int r = 0; r < 3; ++r) {
                double c = x[4 + r], e = 0;
                for (int k = 0; k < 3; ++k) { c += Rm[3 * r + k] * (double)E.nodes[0].c[k]; e += fabs(Rm[3 * r + k]) * (double)E.nodes[0].e[k]; }
                const double g = fabs(ball[4 * a + r] - c) - e;
                if (g > 0) d2 += g * g;
            }
            const double reach = ball[4 * a + 3] + margin;
            if (d2 > reach * reach) dropped[p] = 1;
        }
        if (getenv("CC_DEBUG_TABLES")) {
            size_t nd = 0;
            for (uint8_t d : dropped) nd += d;
            fprintf(stderr, "[tables] reach certificate: %zu environment pairs can never be reached\n", nd);
        }
    }

    // ---- pairs of the batched plan whose relative pose depends on one or two active DOFs, grouped by that DOF set
    struct Cand { int p; int a, ob, j, cls; std::vector<int> dofs; std::vector<double> K; };
    std::vector<Cand> cands;
    for (int p : Q->p_main) {
        if (dropped[p] || Q->table_of_orig[p] >= 0) continue;      // already certified (everywhere / with a table)
        Cand c;
        c.p = p; c.a = Q->quads[4 * p]; c.ob = Q->quads[4 * p + 1]; c.j = Q->quads[4 * p + 2]; c.cls = Q->quads[4 * p + 3];
        if (c.a >= nlb || (c.cls == 1 && c.j >= nlb)) continue;
        const double rb = c.cls == 1 ? rad(c.j) : link_radius(S->bodies[c.ob]->links[c.j]);
        if (!pair_rates(*B, ci, col_of_dof, c.a, c.cls == 1 ? c.j : -1, rad(c.a), rb, c.dofs, c.K)) continue;
        if (c.cls == 1) {
            // the same pair with the roles swapped (a followed from b's frame) is an equally valid bound: keep the tighter one
            std::vector<int> d2; std::vector<double> K2;
            if (pair_rates(*B, ci, col_of_dof, c.j, c.a, rb, rad(c.a), d2, K2) && d2.size() == c.dofs.size()) {
                double s1 = 0, s2 = 0;
                for (double k : c.K) s1 += k;
                for (double k : K2) s2 += k;
                if (s2 < s1) { c.dofs = d2; c.K = K2; }
            }
        }
        // canonical DOF order inside the pair
        if (c.dofs.size() == 2 && c.dofs[0] > c.dofs[1]) { std::swap(c.dofs[0], c.dofs[1]); std::swap(c.K[0], c.K[1]); }
        cands.push_back(c);
    }
    std::sort(cands.begin(), cands.end(), [](const Cand& x, const Cand& y) { return x.dofs < y.dofs; });

    std::vector<ccdev::PairTable>& tables = Q->h_tables;
    std::vector<uint32_t>& bits = Q->h_bits;
    std::vector<int32_t>& table_of_orig = Q->table_of_orig;
    const size_t max_bits_words = (size_t)1 << 20;                  // 4 MB of tables at most
    // One pass over a set of candidates that share their DOF set, on an n0 x n1 grid.  Pairs certified in every cell are
    // marked dropped; with `final` the partly certified ones get their bit table, otherwise they are returned for a finer pass.
    auto grid_pass = [&](const std::vector<size_t>& members, int n0, int n1, bool final, std::vector<size_t>& undecided) -> int {
        const std::vector<int>& dofs = cands[members[0]].dofs;
        const int nd = (int)dofs.size();
        const cc_joint& j0 = *joint_of_dof[dofs[0]];
        const cc_joint* j1 = nd == 2 ? joint_of_dof[dofs[1]] : nullptr;
        const double w0 = (j0.upper - j0.lower) / n0, w1 = j1 ? (j1->upper - j1->lower) / n1 : 0.0;
        if (!(w0 > 0) || (j1 && !(w1 > 0))) return CC_OK;
        // one pairs-query for the whole set: env-class triples keep their order, then the self-class ones
        std::vector<int32_t> triples;
        std::vector<size_t> order;                                  // candidate index of every pair of the group query
        for (int pass = 0; pass < 2; ++pass)
            for (size_t i : members)
                if ((cands[i].cls == 1) == (pass == 1)) {
                    triples.push_back(cands[i].a); triples.push_back(cands[i].cls == 1 ? Q->body : cands[i].ob); triples.push_back(cands[i].j);
                    order.push_back(i);
                }
        double cap = 0;
        std::vector<double> delta(order.size());
        for (size_t k = 0; k < order.size(); ++k) {
            const Cand& c = cands[order[k]];
            delta[k] = 0.5 * w0 * c.K[0] + (j1 ? 0.5 * w1 * c.K[1] : 0.0);
            cap = std::max(cap, delta[k] + margin);
        }
        cc_query* PQ = nullptr;
        int rc = cc_query_create_pairs(S, Q->body, triples.data(), (int)order.size(), &PQ);
        if (rc != CC_OK) return rc;
        const int ncell = n0 * n1;
        std::vector<double> rows((size_t)ncell * B->ndof);
        for (int i0 = 0; i0 < n0; ++i0)
            for (int i1 = 0; i1 < n1; ++i1) {
                double* r = &rows[((size_t)i0 * n1 + i1) * B->ndof];
                for (int d = 0; d < B->ndof; ++d) r[d] = B->dofvals[d];
                r[dofs[0]] = j0.lower + (i0 + 0.5) * w0;
                if (j1) r[dofs[1]] = j1->lower + (i1 + 0.5) * w1;
            }
        if (!PQ->dist_ready) {
            std::vector<uint8_t> all_dynamic(PQ->nq, 1);
            rc = build_plan(PQ, PQ->dist, PQ->p_full, all_dynamic, true);
            if (rc != CC_OK) { cc_query_destroy(PQ); return rc; }
            CC_CUDA(cudaDeviceSynchronize());
            PQ->dist_ready = true;
        }
        const int npd = (int)PQ->dist.pairs.size();
        std::vector<double> dist((size_t)ncell * std::max(1, npd));
        if (npd > 0)
            rc = launch_distance(PQ, PQ->dist, false, rows.data(), ncell, CC_BATCH_ENV | CC_BATCH_SELF, cap * 1.01 + 1e-3, 0.0, 0.0,
                                 dist.data(), nullptr, nullptr, true);
        if (rc != CC_OK) { cc_query_destroy(PQ); return rc; }
        const double two_pi = 6.283185307179586476925286766559;
        auto partial_circle = [&](const cc_joint* j) { return j && j->is_circular && (j->upper - j->lower) < two_pi - 1e-6; };
        for (int dp = 0; dp < npd; ++dp) {
            const size_t k = (size_t)PQ->dist.orig[dp];             // index into `order`
            const Cand& c = cands[order[k]];
            size_t nfree = 0;
            std::vector<uint32_t> tb((ncell + 31) / 32, 0u);
            double dmin = 1e30, slack = 1e30;
            int tight = -1;
            for (int i = 0; i < ncell; ++i) {
                const double d = dist[(size_t)i * npd + dp];
                dmin = std::min(dmin, d);
                if (d > delta[k] + margin) {
                    tb[i >> 5] |= 1u << (i & 31); ++nfree;
                    if (d - delta[k] - margin < slack) { slack = d - delta[k] - margin; tight = i; }
                }
            }
            // what was proven, for cc_query_get_certificates (the adversarial tests aim at the tightest cells)
            cc_certificate rec;
            memset(&rec, 0, sizeof(rec));
            rec.pair = c.p; rec.col0 = col_of_dof[dofs[0]]; rec.col1 = j1 ? col_of_dof[dofs[1]] : -1;
            rec.n0 = n0; rec.n1 = n1; rec.certified_cells = (int32_t)nfree; rec.tightest_cell = tight;
            rec.lo0 = j0.lower; rec.width0 = w0; rec.lo1 = j1 ? j1->lower : 0.0; rec.width1 = w1;
            rec.k0 = c.K[0]; rec.k1 = nd == 2 ? c.K[1] : 0.0; rec.delta = delta[k]; rec.margin = margin;
            rec.min_slack = slack; rec.min_distance = dmin;
            if (getenv("CC_DEBUG_TABLES"))
                fprintf(stderr, "[tables] %s x %s/%s: %d DOF, grid %d x %d, K = %.3f %.3f, delta = %.4f m, min dist = %.4f m, certified cells %.1f%%\n",
                        B->links[c.a].name.c_str(), c.cls == 1 ? B->name.c_str() : S->bodies[c.ob]->name.c_str(),
                        (c.cls == 1 ? B->links[c.j] : S->bodies[c.ob]->links[c.j]).name.c_str(), nd, n0, n1, c.K[0], nd == 2 ? c.K[1] : 0.0,
                        delta[k], dmin, 100.0 * nfree / ncell);
            // never in contact inside the joint limits: the pair can leave the plan — unless a circular DOF covers only part of
            // the circle (SetDOFValues does not clamp those, so values outside the grid exist): then it keeps a table
            if (nfree == (size_t)ncell && !partial_circle(&j0) && !partial_circle(j1)) { dropped[c.p] = 1; rec.dropped = 1; Q->certs.push_back(rec); continue; }
            if (!final) { undecided.push_back(order[k]); continue; }
            if (nfree * 10 < (size_t)ncell || bits.size() + tb.size() > max_bits_words) continue;
            Q->certs.push_back(rec);
            ccdev::PairTable T;
            memset(&T, 0, sizeof(T));
            T.col0 = col_of_dof[dofs[0]]; T.col1 = j1 ? col_of_dof[dofs[1]] : -1;
            T.n0 = n0; T.n1 = n1;
            T.circ0 = !j0.is_circular ? 0 : (partial_circle(&j0) ? 1 : 2);
            T.circ1 = !j1 || !j1->is_circular ? 0 : (partial_circle(j1) ? 1 : 2);
            T.lo0 = j0.lower; T.hi0 = j0.upper; T.inv0 = 1.0 / w0;
            T.lo1 = j1 ? j1->lower : 0.0; T.hi1 = j1 ? j1->upper : 0.0; T.inv1 = j1 ? 1.0 / w1 : 0.0;
            T.bit_base = (uint32_t)bits.size() * 32u;
            bits.insert(bits.end(), tb.begin(), tb.end());
            table_of_orig[c.p] = (int32_t)tables.size();
            tables.push_back(T);
        }
        cc_query_destroy(PQ);
        return CC_OK;
    };
    // coarse grid first (far-apart pairs certify there for a hundredth of the cost), the fine grid only for what is left
    for (size_t g0 = 0; g0 < cands.size();) {
        size_t g1 = g0;
        while (g1 < cands.size() && cands[g1].dofs == cands[g0].dofs) ++g1;
        std::vector<size_t> members, undecided, none;
        for (size_t i = g0; i < g1; ++i) members.push_back(i);
        const bool one = cands[g0].dofs.size() == 1;
        int rc = grid_pass(members, one ? 64 : 32, one ? 1 : 32, false, undecided);
        if (rc != CC_OK) return rc;
        if (!undecided.empty()) {
            std::sort(undecided.begin(), undecided.end());
            rc = grid_pass(undecided, one ? 1024 : 256, one ? 1 : 256, true, none);
            if (rc != CC_OK) return rc;
        }
        g0 = g1;
    }

    // ---- certified-free pairs leave the batched plan
    CC_CUDA(cudaSetDevice(S->device));
    CC_CUDA(cudaDeviceSynchronize());
    std::vector<int> p_keep;
    int newly_dropped = 0;
    for (int p : Q->p_main) {
        if (dropped[p]) { Q->ndropped++; newly_dropped += dropped_before[p] ? 0 : 1; continue; }
        p_keep.push_back(p);
    }
    if (getenv("CC_DEBUG_TABLES"))
        fprintf(stderr, "[tables] plan before: %zu pairs, %d dynamic slots, %d static, %zu fk steps; dropping %d\n", Q->mainl.pairs.size(), Q->mainl.ndyn,
                Q->mainl.nstatic, Q->mainl.steps.size(), Q->ndropped);
    if (p_keep.size() != Q->p_main.size()) {
        std::vector<uint8_t> main_dynamic(Q->nq, 0);
        for (int p : p_keep) {
            const int a = Q->quads[4 * p];
            if (Q->moves[a]) main_dynamic[a] = 1;
            if (Q->quads[4 * p + 3] == 1) { const int b2 = Q->quads[4 * p + 2]; if (Q->moves[b2]) main_dynamic[b2] = 1; }
        }
        int rc = build_plan(Q, Q->mainl, p_keep, main_dynamic, true);
        if (rc != CC_OK) return rc;
        Q->p_main = p_keep;
    }
    if (getenv("CC_DEBUG_TABLES"))
        fprintf(stderr, "[tables] plan after: %zu pairs, %d dynamic slots, %d static, %zu fk steps, %zu tables\n", Q->mainl.pairs.size(), Q->mainl.ndyn,
                Q->mainl.nstatic, Q->mainl.steps.size(), tables.size());
    {
        const int rc = upload_pair_tables(Q);
        if (rc != CC_OK) return rc;
    }
    CC_CUDA(cudaDeviceSynchronize());
    (void)newly_dropped;
    return CC_OK;
}

// Tables cost a few hundred milliseconds to build: they are built once a query has seen enough batched configurations.
extern "C" { static int invalidate_certificates(cc_query* q, const std::vector<uint8_t>* invalid); }

static int maybe_build_tables(cc_query* q, int n)
{
    // the certificates were proven with a margin that includes the triangle-test epsilon in force at the time
    if (q->tables_done && q->cert_tri_eps != q->scene->tri_eps && (q->ntables > 0 || q->ndropped > 0)) {
        int rc = invalidate_certificates(q, nullptr);
        if (rc != CC_OK) return rc;
    }
    if (q->tables_done) return CC_OK;
    static const long long after = getenv("CC_TUNE_TABLES_AFTER") ? atoll(getenv("CC_TUNE_TABLES_AFTER")) : 16384;
    q->fk_configs_seen += (uint64_t)n;
    if ((long long)q->fk_configs_seen < after) return CC_OK;
    // build_pair_tables proves separation from the LIVE scene state (link poses, inactive DOF values, a nested pairs query);
    // the batched kernels check against the snapshot the query captured at cc_query_create / cc_query_refresh.  The two are the
    // same only while no body has been moved since: otherwise wait for the caller's cc_query_refresh (certificates already in
    // force stay valid: they were proven on the snapshot the query still uses, and a refresh drops them).
    if (q->pose_epoch != q->scene->pose_epoch) return CC_OK;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = build_pair_tables(q);
    q->cert_build_ms = 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

