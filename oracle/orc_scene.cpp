// orc_scene.cpp -- CPU oracle: BVH2 + brute-force ray queries.  TEST INFRASTRUCTURE ONLY.
#include "orc_scene.h"
#include <algorithm>
#include <cfloat>

namespace orc {

// Intersection contract, identical operation order to csrc/traverse.cuh (see the header comment
// there).  Reference behaviour restated: closest hit of optixTrace with
// OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES (cuProg.h:402,427,452) where only emitter quads are
// cullable (sutil/Scene.cpp:1030,1085 sets DISABLE_TRIANGLE_FACE_CULLING for doubleSided
// materials, scene_shift.cpp:68; light materials keep the default single-sided MaterialData.h:114).
// Front face = counter-clockwise seen from the ray origin  <=>  det > 0 below.
bool orc_tri_test(const Tri& tr, f3 o, f3 d, float tmin, bool cull_back, float& t, float& u, float& v) {
    const f3 pvec = c_cross(d, tr.e2);
    const float det = c_dot(tr.e1, pvec);
    const bool single = cull_back && tr.light >= 0;
    if (single ? !(det > 0.0f) : !(det != 0.0f)) return false;
    const float inv = 1.0f / det;
    const f3 tvec = o - tr.v0;
    u = c_dot(tvec, pvec) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const f3 qvec = c_cross(tvec, tr.e1);
    v = c_dot(d, qvec) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = c_dot(tr.e2, qvec) * inv;
    if (!(t > tmin)) return false;
    return true;
}

void Scene::build_bvh() {
    const int n = (int)tris.size();
    order.resize(n);
    std::vector<f3> lo(n), hi(n), cen(n);
    float maxabs = 0.f;
    for (int i = 0; i < n; i++) {
        order[i] = i;
        const Tri& t = tris[i];
        lo[i] = mk3(std::min(t.v0.x, std::min(t.v1.x, t.v2.x)), std::min(t.v0.y, std::min(t.v1.y, t.v2.y)), std::min(t.v0.z, std::min(t.v1.z, t.v2.z)));
        hi[i] = mk3(std::max(t.v0.x, std::max(t.v1.x, t.v2.x)), std::max(t.v0.y, std::max(t.v1.y, t.v2.y)), std::max(t.v0.z, std::max(t.v1.z, t.v2.z)));
        cen[i] = (lo[i] + hi[i]) * 0.5f;
        maxabs = std::max(maxabs, std::max(std::max(std::fabs(lo[i].x), std::fabs(lo[i].y)), std::fabs(lo[i].z)));
        maxabs = std::max(maxabs, std::max(std::max(std::fabs(hi[i].x), std::fabs(hi[i].y)), std::fabs(hi[i].z)));
    }
    pad = std::max(maxabs, 1e-20f) * (1.0f / 65536.0f);   // generous: the oracle only needs to be conservative
    nodes.clear();
    nodes.reserve(2 * n);
    struct Job { int node, first, count; };
    std::vector<Job> stack;
    nodes.push_back(Bvh2Node{});
    stack.push_back(Job{0, 0, n});
    while (!stack.empty()) {
        const Job j = stack.back();
        stack.pop_back();
        f3 blo = mk3(FLT_MAX), bhi = mk3(-FLT_MAX), clo = mk3(FLT_MAX), chi = mk3(-FLT_MAX);
        for (int k = j.first; k < j.first + j.count; k++) {
            const int p = order[k];
            blo = mk3(std::min(blo.x, lo[p].x), std::min(blo.y, lo[p].y), std::min(blo.z, lo[p].z));
            bhi = mk3(std::max(bhi.x, hi[p].x), std::max(bhi.y, hi[p].y), std::max(bhi.z, hi[p].z));
            clo = mk3(std::min(clo.x, cen[p].x), std::min(clo.y, cen[p].y), std::min(clo.z, cen[p].z));
            chi = mk3(std::max(chi.x, cen[p].x), std::max(chi.y, cen[p].y), std::max(chi.z, cen[p].z));
        }
        Bvh2Node nd;
        nd.lo = blo - mk3(pad);
        nd.hi = bhi + mk3(pad);
        nd.left = nd.right = -1;
        nd.first = j.first;
        nd.count = j.count;
        if (j.count > 4) {
            const f3 ext = chi - clo;
            int axis = 0;
            if (ext.y > ext.x) axis = 1;
            if (ext.z > (axis == 0 ? ext.x : ext.y)) axis = 2;
            const int mid = j.first + j.count / 2;
            auto key = [&](int p) { return axis == 0 ? cen[p].x : (axis == 1 ? cen[p].y : cen[p].z); };
            std::nth_element(order.begin() + j.first, order.begin() + mid, order.begin() + j.first + j.count,
                             [&](int a, int b) { return key(a) < key(b) || (key(a) == key(b) && a < b); });
            nd.left = (int)nodes.size();
            nodes.push_back(Bvh2Node{});
            nd.right = (int)nodes.size();
            nodes.push_back(Bvh2Node{});
            nd.count = 0;
            stack.push_back(Job{nd.left, j.first, mid - j.first});
            stack.push_back(Job{nd.right, mid, j.first + j.count - mid});
        }
        nodes[j.node] = nd;
    }
}

static inline bool slab(const Bvh2Node& nd, f3 o, f3 id, float tmin, float tmax) {
    float t0 = tmin, t1 = tmax;
    const float ax = (nd.lo.x - o.x) * id.x, bx = (nd.hi.x - o.x) * id.x;
    const float ay = (nd.lo.y - o.y) * id.y, by = (nd.hi.y - o.y) * id.y;
    const float az = (nd.lo.z - o.z) * id.z, bz = (nd.hi.z - o.z) * id.z;
    t0 = std::max(t0, std::min(ax, bx)); t1 = std::min(t1, std::max(ax, bx));
    t0 = std::max(t0, std::min(ay, by)); t1 = std::min(t1, std::max(ay, by));
    t0 = std::max(t0, std::min(az, bz)); t1 = std::min(t1, std::max(az, bz));
    return t0 <= t1 * 1.0000004f + 1e-30f || !(t0 == t0) || !(t1 == t1);   // NaN -> visit (conservative)
}

static inline f3 safe_inv(f3 d) {
    const float eps = 1.0e-24f;
    return mk3(1.0f / (std::fabs(d.x) > eps ? d.x : std::copysign(eps, d.x)),
               1.0f / (std::fabs(d.y) > eps ? d.y : std::copysign(eps, d.y)),
               1.0f / (std::fabs(d.z) > eps ? d.z : std::copysign(eps, d.z)));
}

bool Scene::closest(f3 o, f3 d, float tmin, float tmax, bool cull_back, Hit& h, bool brute) const {
    float best_t = tmax;
    int best = -1;
    float bu = 0.f, bv = 0.f;
    auto test = [&](int p) {
        float t, u, v;
        if (!orc_tri_test(tris[p], o, d, tmin, cull_back, t, u, v)) return;
        if (t < best_t || (t == best_t && p < best)) { best_t = t; best = p; bu = u; bv = v; }
    };
    if (brute || nodes.empty()) {
        for (int p = 0; p < (int)tris.size(); p++) test(p);
    } else {
        const f3 id = safe_inv(d);
        int stack[128];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Bvh2Node& nd = nodes[stack[--sp]];
            if (!slab(nd, o, id, tmin, best_t)) continue;
            if (nd.left < 0) {
                for (int k = nd.first; k < nd.first + nd.count; k++) test(order[k]);
            } else {
                stack[sp++] = nd.left;
                stack[sp++] = nd.right;
            }
        }
    }
    h.prim = best;
    h.t = best >= 0 ? best_t : 0.f;
    h.u = bu;
    h.v = bv;
    return best >= 0;
}

bool Scene::occluded(f3 o, f3 d, float tmin, float tmax, bool brute) const {
    auto test = [&](int p) {
        float t, u, v;
        return orc_tri_test(tris[p], o, d, tmin, false, t, u, v) && t < tmax;
    };
    if (brute || nodes.empty()) {
        for (int p = 0; p < (int)tris.size(); p++)
            if (test(p)) return true;
        return false;
    }
    const f3 id = safe_inv(d);
    int stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
        const Bvh2Node& nd = nodes[stack[--sp]];
        if (!slab(nd, o, id, tmin, tmax)) continue;
        if (nd.left < 0) {
            for (int k = nd.first; k < nd.first + nd.count; k++)
                if (test(order[k])) return true;
        } else {
            stack[sp++] = nd.left;
            stack[sp++] = nd.right;
        }
    }
    return false;
}

}  // namespace orc
