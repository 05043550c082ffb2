// tree_host.cu -- spc_build_tree: the subspace classification trees built on the host, without a context or a device.
//
// ABI counterpart of classTree::buildTreeBaseOnExistSample::operator()(samples, K, labelBias) (decisionTree/classTree_host.h:302-431),
// the one host-side call of the reference's preprocessing seam (optixPathTracer.cpp:563-567).  The product's default is the device-side
// build (tree_build.cu); this entry point serves callers that hold the weighted points on the host, and it is the host twin of that
// builder: the same data layout (one sample order per level, a node is a range of it) and the same bookkeeping, run serially.
//
// What has to hold for the node array to equal the reference's bit for bit (tests/golden/tree.npz, tests/test_oracle_vs_ref.py):
//   * centres: sample i is a centre when the running weight passes weight_sum / K (:313-322); labels: nearest centre under
//     |dp|^2 + s2 * (1 - n.n'), first minimum wins (classTree_common.h:82-90; s2 = largest per-axis position variance, :287-301);
//   * nodes are created breadth first, eight at a time, and node i is split while it is impure, shallower than 15 and the weighted
//     accuracy c_w of the whole tree is below 0.99, c_w being updated after every split (:344-372);
//   * a split is by position on even depth or once four normal splits lie above the node, else by normal (DIR_JUDGE 0); its centre is
//     the nearest same-type ancestor's centre moved by the level's half extent towards the octant the node descends from (:103-211);
//   * every fp32 sum runs over a node's samples in their original relative order: the partition into octants is stable, a child's
//     weight is the in-order sum of its samples' weights, and the majority vote keeps a running maximum of the per-label sums, so of
//     two labels with equal totals the one whose sum got there first wins (:243-284).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <thread>
#include <vector>
#include "common.cuh"

namespace {

constexpr int   kDepthLimit = 15;      // classTree_host.h:344
constexpr float kAccuracy = 0.99f;

struct F3 {
    float x, y, z;
};
inline F3 f3(const spc_float3& v) { return F3{v.x, v.y, v.z}; }
inline F3 halved(F3 a) {   // the reference divides vectors by multiplying with the reciprocal (sutil/vec_math.h:483-487)
    const float r = 1.0f / 2.0f;
    return F3{a.x * r, a.y * r, a.z * r};
}

// samples as the builder needs them: the two keys an octant test can read, weight (normalised), label
struct Points {
    std::vector<F3>    key[2];   // [0] position, [1] normal
    std::vector<float> w;
    std::vector<int>   label;
};

// ---- step 1 + 2: centres and nearest-centre labels ---------------------------------------------------------------------------
struct Centres {
    std::vector<F3> pos, nrm, dir;
};

float position_spread(const spc_divide_weight* s, int n) {   // get_position_variance: mean and variance accumulate pre-divided terms
    const float inv_n = 1.0f / (float)n, inv_nm1 = 1.0f / (float)(n - 1);
    F3 mean{0.f, 0.f, 0.f};
    for (int i = 0; i < n; i++) {
        mean.x = mean.x + s[i].position.x * inv_n;
        mean.y = mean.y + s[i].position.y * inv_n;
        mean.z = mean.z + s[i].position.z * inv_n;
    }
    F3 var{0.f, 0.f, 0.f};
    for (int i = 0; i < n; i++) {
        const float dx = mean.x - s[i].position.x, dy = mean.y - s[i].position.y, dz = mean.z - s[i].position.z;
        var.x = var.x + (dx * dx) * inv_nm1;
        var.y = var.y + (dy * dy) * inv_nm1;
        var.z = var.z + (dz * dz) * inv_nm1;
    }
    return fmaxf(var.x, fmaxf(var.y, var.z));
}

Centres pick_centres(const spc_divide_weight* s, int n, int K) {
    float total = 0;
    for (int i = 0; i < n; i++) total += s[i].weight;
    const float quota = total / K;
    Centres c;
    float run = 0;
    for (int i = 0; i < n; i++) {
        run += s[i].weight;
        if (run > quota) {
            run -= quota;
            c.pos.push_back(f3(s[i].position));
            c.nrm.push_back(f3(s[i].normal));
            c.dir.push_back(f3(s[i].dir));
        }
    }
    return c;
}

void label_points(const spc_divide_weight* s, int n, const Centres& c, float spread, int label_bias, std::vector<int>& label) {
    const int nc = (int)c.pos.size();
    auto work = [&](int b, int e) {
        for (int i = b; i < e; i++) {
            const F3 p = f3(s[i].position), pn = f3(s[i].normal), pd = f3(s[i].dir);
            float best = FLT_MAX;
            int arg = 0;     // (no centre closer than FLT_MAX: label 0, without the bias -- as the reference)
            for (int k = 0; k < nc; k++) {
                const float dx = c.pos[k].x - p.x, dy = c.pos[k].y - p.y, dz = c.pos[k].z - p.z;
                const float dist2 = dx * dx + dy * dy + dz * dz;
                const float along_d = pd.x * c.dir[k].x + pd.y * c.dir[k].y + pd.z * c.dir[k].z;
                const float along_n = pn.x * c.nrm[k].x + pn.y * c.nrm[k].y + pn.z * c.nrm[k].z;
                // divide_weight::d with k = DIR_JUDGE = 0: the direction term is multiplied by zero but still evaluated (NaN stays NaN)
                const float d = dist2 + spread * ((1 - along_n) + (1 - along_d) * 0.0f);
                if (d < best) {
                    best = d;
                    arg = k + label_bias;
                }
            }
            label[i] = arg;
        }
    };
    const int T = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const int per = (n + T - 1) / T;
    std::vector<std::thread> pool;
    for (int t = 0; t < T; t++) pool.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
    for (auto& th : pool) th.join();
}

// ---- step 3: the octree ------------------------------------------------------------------------------------------------------
struct Build {
    // per node, parallel to `out`
    struct Info {
        int   begin, end;        // its samples: [begin, end) of its level's order
        int   depth, n_pos, n_nrm;   // splits above it: all, by position, by normal
        float weight, majority;  // in-order sum of its samples' weights; weight of its majority label
        F3    anchor[2];         // centre of the nearest ancestor split by position / by normal ...
        int   octant[2];         // ... and the octant of that ancestor this node lies in
    };
    std::vector<spc_tree_node> out;
    std::vector<Info>          info;
    const Points&              pts;
    std::vector<float>         tally;       // per-label running sums of one vote; valid where seen[label] == vote_no
    std::vector<int>           seen;
    int                        vote_no = 0;

    Build(const Points& p, int n_labels) : pts(p), tally((size_t)n_labels, 0.f), seen((size_t)n_labels, 0) {}

    int add_node(int label) {
        spc_tree_node nd;
        memset(&nd, 0, sizeof(nd));
        nd.leaf = 1;
        nd.label = label;
        out.push_back(nd);
        info.push_back(Info{});
        return (int)out.size() - 1;
    }

    // majority label of node `id` over order[begin, end): leaves label and `majority` as the reference's colour pass does
    void vote(int id, const std::vector<int>& order) {
        Info& f = info[id];
        if (f.begin == f.end) {     // empty: keeps the label it was created with, no correct weight
            f.majority = 0.0f;
            return;
        }
        const int first = pts.label[order[f.begin]];
        bool pure = true;
        for (int k = f.begin + 1; k < f.end && pure; k++) pure = pts.label[order[k]] == first;
        if (pure) {
            out[id].label = first;
            f.majority = f.weight;
            return;
        }
        float top = 0.0f;
        int   top_label = first;
        vote_no++;
        for (int k = f.begin; k < f.end; k++) {
            const int l = pts.label[order[k]];
            if (seen[l] != vote_no) {
                seen[l] = vote_no;
                tally[l] = 0.f;
            }
            const float sum = (tally[l] += pts.w[order[k]]);
            if (top < sum) {
                top = sum;
                top_label = l;
            }
        }
        out[id].label = top_label;
        f.majority = top;
    }

    // splits node `id` (samples in `cur`), appends its children's samples to `next`; returns the children's correct weight
    float split(int id, const std::vector<int>& cur, std::vector<int>& next, const std::vector<F3>& extent_pos, const std::vector<F3>& extent_nrm) {
        const Info f = info[id];
        const int  type = (f.depth % 2 == 0 || f.n_nrm > 3) ? 0 : 1;
        F3 mid;
        if (type == 1 && f.n_nrm == 0) mid = F3{0.f, 0.f, 0.f};                    // the first normal split is about the origin
        else if (f.n_pos == 0) mid = f3(out[id].mid);                               // the root: centre of the bounding box
        else {
            const F3 h = type == 0 ? extent_pos[f.n_pos + 1] : extent_nrm[f.n_nrm + 1];
            const int o = f.octant[type];
            mid = F3{f.anchor[type].x + ((o & 1) ? h.x : -h.x), f.anchor[type].y + ((o & 2) ? h.y : -h.y), f.anchor[type].z + ((o & 4) ? h.z : -h.z)};
        }
        out[id].leaf = 0;
        out[id].type = type;
        out[id].mid = spc_float3{mid.x, mid.y, mid.z};
        // stable partition of the node's samples into the eight octants
        const std::vector<F3>& key = pts.key[type];
        int count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        auto octant_of = [&](int s) { return (key[s].x > mid.x ? 1 : 0) + (key[s].y > mid.y ? 2 : 0) + (key[s].z > mid.z ? 4 : 0); };
        for (int k = f.begin; k < f.end; k++) count[octant_of(cur[k])]++;
        int start[8], at = (int)next.size();
        for (int o = 0; o < 8; o++) {
            start[o] = at;
            at += count[o];
        }
        next.resize((size_t)at);
        float wsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        {
            int fill[8];
            for (int o = 0; o < 8; o++) fill[o] = start[o];
            for (int k = f.begin; k < f.end; k++) {
                const int s = cur[k], o = octant_of(s);
                next[fill[o]++] = s;
                wsum[o] += pts.w[s];
            }
        }
        float correct = 0.0f;
        for (int o = 0; o < 8; o++) {
            const int ch = add_node(out[id].label);
            out[id].child[o] = ch;
            Info& c = info[ch];
            c.begin = start[o];
            c.end = start[o] + count[o];
            c.depth = f.depth + 1;
            c.n_pos = f.n_pos + (type == 0);
            c.n_nrm = f.n_nrm + (type == 1);
            c.weight = wsum[o];
            c.anchor[type] = mid;
            c.octant[type] = o;
            c.anchor[1 - type] = f.anchor[1 - type];
            c.octant[1 - type] = f.octant[1 - type];
            vote(ch, next);
            correct += c.majority;
        }
        return correct;
    }

    void run(int n) {
        // extents of the position / normal cells per split depth: the bounding box (sic: the maximum starts from FLT_MIN, the smallest
        // positive float, classTree_host.h:99-100) and the cube [-1, 1]^3, halved per level
        F3 lo{FLT_MAX, FLT_MAX, FLT_MAX}, hi{FLT_MIN, FLT_MIN, FLT_MIN};
        for (int i = 0; i < n; i++) {
            const F3 p = pts.key[0][i];
            lo = F3{fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z)};
            hi = F3{fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z)};
        }
        std::vector<F3> extent_pos, extent_nrm;
        F3 e{hi.x - lo.x, hi.y - lo.y, hi.z - lo.z};
        for (int i = 0; i < kDepthLimit + 10; i++, e = halved(e)) extent_pos.push_back(e);
        e = F3{2.0f, 2.0f, 2.0f};
        for (int i = 0; i < 15; i++, e = halved(e)) extent_nrm.push_back(e);

        std::vector<int> cur((size_t)n), next;
        for (int i = 0; i < n; i++) cur[i] = i;
        const int root = add_node(0);
        info[root].begin = 0;
        info[root].end = n;
        info[root].weight = 1;
        const F3 centre = halved(F3{hi.x + lo.x, hi.y + lo.y, hi.z + lo.z});
        out[root].mid = spc_float3{centre.x, centre.y, centre.z};
        vote(root, cur);
        float accuracy = info[root].majority;
        // level by level; inside a level in node order, which is the reference's breadth-first order
        int level_begin = 0;
        while (level_begin < (int)out.size()) {
            const int level_end = (int)out.size();
            next.clear();
            for (int id = level_begin; id < level_end; id++) {
                const Info f = info[id];     // (by value: a split appends to `info`)
                if (f.begin < f.end && f.majority < f.weight && f.depth < kDepthLimit && kAccuracy > accuracy) {
                    accuracy -= f.majority;
                    accuracy += split(id, cur, next, extent_pos, extent_nrm);
                }
            }
            cur.swap(next);
            level_begin = level_end;
        }
    }
};

}  // namespace

extern "C" {

// Returns the number of nodes (also when it exceeds `cap`, in which case nothing is written), or a negative spc_status.
int spc_build_tree(const spc_divide_weight* samples, int n, int K, int label_bias, spc_tree_node* out, int cap, int* max_label) {
    if (!samples || n < 2 || K < 1 || !out) {
        spc::set_error("spc_build_tree: bad arguments");
        return SPC_ERR_INVALID;
    }
    const float spread = position_spread(samples, n);
    const Centres centres = pick_centres(samples, n, K);
    Points pts;
    pts.label.resize((size_t)n);
    label_points(samples, n, centres, spread, label_bias, pts.label);
    // para_initial (:213-241): weights normalised by their in-order sum
    pts.key[0].resize((size_t)n);
    pts.key[1].resize((size_t)n);
    pts.w.resize((size_t)n);
    float norm = 0.0f;
    for (int i = 0; i < n; i++) norm += samples[i].weight;
    for (int i = 0; i < n; i++) {
        pts.key[0][i] = f3(samples[i].position);
        pts.key[1][i] = f3(samples[i].normal);
        pts.w[i] = samples[i].weight / norm;
    }
    Build b(pts, (int)centres.pos.size() + label_bias + 1);
    b.run(n);
    int top = 0;
    for (const spc_tree_node& nd : b.out) top = std::max(top, nd.label);
    if (max_label) *max_label = top;
    const int size = (int)b.out.size();
    if (size <= cap) std::copy(b.out.begin(), b.out.end(), out);
    return size;
}

}  // extern "C"
