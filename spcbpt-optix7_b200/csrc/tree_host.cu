// tree_host.cu -- host-side builder of the subspace classification trees.
// Stands in for classTree::buildTreeBaseOnExistSample::operator()(samples, K, labelBias)
// (decisionTree/classTree_host.h:302-431), which the reference also runs on the host, once, during
// preprocessing (optixPathTracer.cpp:563-567).  Same algorithm and the same fp32 evaluation order, so the
// produced tree_node array equals the reference's on the same samples (tests/golden/tree.npz):
//   1. centres: every time the running sample weight passes sum/K (classTree_host.h:313-322)
//   2. label  : nearest centre under |dp|^2 + s2 * (1 - n.n')  (classTree_common.h:82-90; s2 = largest
//               per-axis position variance, classTree_host.h:287-301); this O(N*K) loop is threaded
//   3. octree : BFS over nodes; a node is split (8 children; position split on even depth or once 4 normal
//               splits were made, else normal split) while it is impure, depth < 15 and the weighted accuracy
//               of the whole tree is below 0.99 (classTree_host.h:103-211, 243-284, 344-372)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <thread>
#include <vector>
#include "common.cuh"

namespace {

struct V3 {
    float x, y, z;
};
inline V3 v3(float a, float b, float c) { return V3{a, b, c}; }
inline V3 ld(const spc_float3& f) { return V3{f.x, f.y, f.z}; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 div_s(V3 a, float s) {   // sutil/vec_math.h:483-487: multiply by the reciprocal
    const float inv = 1.0f / s;
    return a * inv;
}
inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct Sample {
    V3 position, dir, normal;
    float weight;
    int label;
};

struct Node {
    spc_tree_node n;
    std::vector<Sample> v;
    int depth = 0;
    float weight = 0.f;
    float correct_weight = 0.f;
    int father = 0;
    int position_depth = 0, normal_depth = 0, dir_depth = 0;
    Node() {
        memset(&n, 0, sizeof(n));
        n.leaf = 1;
        n.label = 0;
        n.type = 0;
    }
};

struct Builder {
    std::vector<Node> v;
    std::vector<V3> block_size, direction_block_size;
    V3 bbox_min = v3(FLT_MAX, FLT_MAX, FLT_MAX);
    V3 bbox_max = v3(FLT_MIN, FLT_MIN, FLT_MIN);   // sic: FLT_MIN (smallest positive), classTree_host.h:99-100
    int n_labels = 0;

    int child_of(const Node& nd, const Sample& s) const {   // tree_node::operator(), classTree_common.h:20-36
        const V3 q = nd.n.type == 0 ? s.position : (nd.n.type == 1 ? s.normal : s.dir);
        int ind = 0;
        ind += q.x > nd.n.mid.x ? 1 : 0;
        ind += q.y > nd.n.mid.y ? 2 : 0;
        ind += q.z > nd.n.mid.z ? 4 : 0;
        return nd.n.child[ind];
    }

    void color(int id) {   // classTree_host.h:243-284
        Node& t = v[id];
        if (t.v.empty()) {
            t.correct_weight = 0.0f;
            return;
        }
        bool need_split = false;
        t.n.label = t.v[0].label;
        for (size_t i = 0; i < t.v.size(); i++)
            if (t.v[i].label != t.n.label) {
                need_split = true;
                break;
            }
        if (need_split) {
            std::vector<float> weights((size_t)n_labels, 0.f);
            float max_weight = 0.0f;
            int max_weight_id = t.n.label;
            for (size_t i = 0; i < t.v.size(); i++) {
                weights[t.v[i].label] += t.v[i].weight;
                if (max_weight < weights[t.v[i].label]) {
                    max_weight = weights[t.v[i].label];
                    max_weight_id = t.v[i].label;
                }
            }
            t.n.label = max_weight_id;
            t.correct_weight = max_weight;
        } else {
            t.correct_weight = t.weight;
        }
    }

    float split(int id) {   // classTree_host.h:103-211
        const int split_type = (v[id].depth % 2 == 0 || v[id].normal_depth > 3) ? 0 : 1;   // DIR_JUDGE 0: never a direction split
        const int back = (int)v.size();
        v[id].n.leaf = 0;
        V3 inch;
        if (split_type == 0) inch = block_size[v[id].position_depth + 1];
        else inch = direction_block_size[v[id].normal_depth + 1];
        V3 mid;
        if (v[id].normal_depth == 0 && split_type == 1) {
            mid = v3(0.f, 0.f, 0.f);
        } else if (v[id].position_depth == 0) {
            mid = ld(v[id].n.mid);
        } else {
            int L_id = id;
            int t_id = v[id].father;
            while (t_id != 0 && v[t_id].n.type != split_type) {
                L_id = t_id;
                t_id = v[t_id].father;
            }
            mid = ld(v[t_id].n.mid);
            int c = 0;
            for (; c < 8; c++)
                if (v[t_id].n.child[c] == L_id) break;
            const V3 delta = v3((c >> 0) % 2 == 0 ? -inch.x : inch.x, (c >> 1) % 2 == 0 ? -inch.y : inch.y, (c >> 2) % 2 == 0 ? -inch.z : inch.z);
            mid = mid + delta;
        }
        v[id].n.mid = spc_float3{mid.x, mid.y, mid.z};
        v[id].n.type = split_type;
        for (int i = 0; i < 8; i++) {
            v[id].n.child[i] = back + i;
            v.push_back(Node());
            Node& c = v.back();
            c.father = id;
            c.depth = v[id].depth + 1;
            c.n.label = v[id].n.label;
            c.position_depth = v[id].position_depth + (split_type == 0);
            c.normal_depth = v[id].normal_depth + (split_type == 1);
            c.dir_depth = v[id].dir_depth;
        }
        for (size_t k = 0; k < v[id].v.size(); k++) {
            const Sample& s = v[id].v[k];
            Node& c = v[child_of(v[id], s)];
            c.v.push_back(s);
            c.weight += s.weight;
        }
        float n_correct_weight = 0.0f;
        for (int i = 0; i < 8; i++) {
            color(v[id].n.child[i]);
            n_correct_weight += v[v[id].n.child[i]].correct_weight;
        }
        v[id].weight = 0;
        v[id].v.clear();
        v[id].v.shrink_to_fit();
        return n_correct_weight;
    }

    void run(std::vector<Sample>& samples, float threshold, int max_depth, int* max_label_out) {   // classTree_host.h:344-372
        // para_initial (:213-241)
        float unnorm = 0.0f;
        for (auto& p : samples) {
            unnorm += p.weight;
            bbox_min = v3(fminf(bbox_min.x, p.position.x), fminf(bbox_min.y, p.position.y), fminf(bbox_min.z, p.position.z));
            bbox_max = v3(fmaxf(bbox_max.x, p.position.x), fmaxf(bbox_max.y, p.position.y), fmaxf(bbox_max.z, p.position.z));
        }
        for (auto& p : samples) p.weight /= unnorm;
        V3 bb = bbox_max - bbox_min;
        for (int i = 0; i < max_depth + 10; i++) {
            block_size.push_back(bb);
            bb = div_s(bb, 2.0f);
        }
        V3 db = v3(2.0f, 2.0f, 2.0f);
        for (int i = 0; i < 15; i++) {
            direction_block_size.push_back(db);
            db = div_s(db, 2.0f);
        }
        v.push_back(Node());
        v[0].v = samples;
        v[0].weight = 1;
        const V3 m = div_s(bbox_max + bbox_min, 2.0f);
        v[0].n.mid = spc_float3{m.x, m.y, m.z};
        color(0);
        float c_w = v[0].correct_weight;
        int max_label = 0;
        for (size_t i = 0; i < v.size(); i++) {
            max_label = std::max(v[i].n.label, max_label);
            if (!v[i].v.empty() && v[i].correct_weight < v[i].weight && v[i].depth < max_depth && threshold > c_w) {
                c_w -= v[i].correct_weight;
                c_w += split((int)i);
            }
        }
        if (max_label_out) *max_label_out = max_label;
    }
};

}  // namespace

extern "C" {

// classTree::buildTreeBaseOnExistSample()(samples, subspaceSize, labelBias), classTree_host.h:302-343.
// Returns the number of nodes (also when it exceeds `cap`, in which case nothing is written), or a negative
// spc_status.  Pure host code: no context, no device needed (the reference's tree build is host code too).
int spc_build_tree(const spc_divide_weight* samples, int n, int K, int label_bias, spc_tree_node* out, int cap, int* max_label) {
    if (!samples || n < 2 || K < 1 || !out) {
        spc::set_error("spc_build_tree: bad arguments");
        return SPC_ERR_INVALID;
    }
    // get_position_variance (classTree_host.h:287-301)
    const float it = (float)n;
    V3 mean = v3(0.f, 0.f, 0.f);
    for (int i = 0; i < n; i++) mean = mean + div_s(ld(samples[i].position), it);
    V3 var = v3(0.f, 0.f, 0.f);
    const float itm1 = (float)(n - 1);
    for (int i = 0; i < n; i++) {
        const V3 diff = mean - ld(samples[i].position);
        var = var + div_s(diff * diff, itm1);
    }
    const float diversity2 = fmaxf(var.x, fmaxf(var.y, var.z));
    float weight_sum = 0;
    for (int i = 0; i < n; i++) weight_sum += samples[i].weight;
    std::vector<int> centers;
    float acc = 0;
    for (int i = 0; i < n; i++) {
        acc += samples[i].weight;
        if (acc > weight_sum / K) {
            acc -= weight_sum / K;
            centers.push_back(i);
        }
    }
    std::vector<Sample> labeled((size_t)n);
    const int nc = (int)centers.size();
    auto label_range = [&](int b, int e) {
        for (int i = b; i < e; i++) {
            const spc_divide_weight& p = samples[i];
            float min_distance = FLT_MAX;
            int id = 0;
            const V3 pp = ld(p.position), pn = ld(p.normal), pd = ld(p.dir);
            for (int c = 0; c < nc; c++) {
                const spc_divide_weight& a = samples[centers[c]];
                // divide_weight::d (classTree_common.h:82-90), k = DIR_JUDGE = 0
                const V3 diff = ld(a.position) - pp;
                const float d_a = dot3(diff, diff);
                const float diff_direction = dot3(pd, ld(a.dir));
                const float diff_normal = dot3(pn, ld(a.normal));
                const float d = d_a + diversity2 * ((1 - diff_normal) + (1 - diff_direction) * 0.0f);
                if (d < min_distance) {
                    min_distance = d;
                    id = c + label_bias;
                }
            }
            labeled[i] = Sample{pp, pd, pn, p.weight, id};
        }
    };
    {
        const int T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        const int per = (n + T - 1) / T;
        for (int t = 0; t < T; t++) pool.emplace_back(label_range, std::min(n, t * per), std::min(n, (t + 1) * per));
        for (auto& th : pool) th.join();
    }
    Builder b;
    b.n_labels = nc + label_bias + 1;
    int ml = 0;
    b.run(labeled, 0.99f, 15, &ml);
    if (max_label) *max_label = ml;
    const int size = (int)b.v.size();
    if (size <= cap)
        for (int i = 0; i < size; i++) out[i] = b.v[i].n;
    return size;
}

}  // extern "C"
