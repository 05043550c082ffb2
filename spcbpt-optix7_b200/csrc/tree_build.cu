// tree_build.cu -- device-side build of the subspace classification trees.
//
// Stands in for classTree::buildTreeBaseOnExistSample::operator()(samples, K, labelBias) (decisionTree/classTree_host.h:302-431),
// which the reference runs on the host after copying the weighted sample points out of the training set.  Here the points stay in
// HBM: nearest-centre labelling (N x K distances, centres staged through shared memory), then a LEVEL-SYNCHRONOUS octree build.
// The result is bit-equal to the reference's tree (tests/golden/tree.npz, tests/test_tree_build_gpu.py):
//
//   * The reference builds breadth first: node i is split while it is impure, shallower than 15 and the weighted accuracy of the whole
//     tree, c_w, is below 0.99; c_w is updated after every split (classTree_host.h:344-372).  c_w never decreases, so the nodes that are
//     split are a PREFIX (in index order) of the eligible nodes of each level.  All eligible nodes of a level are therefore split
//     speculatively in parallel (children, their weights and majority labels), and one serial pass over the level's nodes then replays
//     the reference's bookkeeping -- which splits count, the children's indices, c_w -- in the reference's own fp32 order.
//   * Every fp32 sum the reference forms in a serial loop is formed in the same order here: samples are kept in their original
//     relative order inside every node (stable sorts), a child's weight and the per-label weights of its majority vote are summed
//     by one thread walking the child's samples in that order (per-label: after a second stable sort by label).
//   * The majority vote keeps a running maximum over the samples (classTree_host.h:262-273): on a tie of two labels' final weights
//     the label whose running sum reached that value FIRST wins; the position where each label's sum last changed is tracked.
//
// One-off stage (two trees per training run, ~10^5 samples): clarity over speed; cub's stable radix sort is the only library call.
#include <cfloat>
#include <cub/cub.cuh>
#include <vector>
#include "common.cuh"

namespace spc {

namespace {

constexpr int kMaxTreeDepth = 15;        // classTree_host.h:344 (max_depth)
constexpr float kAccuracy = 0.99f;       // threshold

struct TbNode {                          // build-time record of a node; `out` is what the caller gets
    int   father, depth, pdepth, ndepth, slot;   // slot: which child of its father
    int   begin, end;                    // its samples: positions [begin, end) of the level's sample order
    float weight, cw;                    // sum of its samples' weights (in order); weight of its majority label
    int   cand;                          // rank among the level's split candidates, or -1
};

__device__ __forceinline__ float3 ld3(const spc_float3& v) { return make_float3(v.x, v.y, v.z); }

// ---- serial sums, fed in parallel ---------------------------------------------------------------------------------------------
// The reference forms these sums in serial host loops, and fp32 addition is not associative: to get the same bits the adds must
// happen in the same order.  One thread does the adds; the rest of its block only stages the operands through shared memory in
// coalesced tiles, so the serial chain runs at shared-memory speed (~10 cycles per element) instead of one dependent global load
// per element.
constexpr int kTbBlock = 256, kTbTile = 2048;

// prologue (classTree_host.h:287-322): block 0 -> weight_sum and the centre list; blocks 1..3 -> mean and variance of x / y / z
__global__ void __launch_bounds__(kTbBlock) k_tb_prologue(const spc_divide_weight* __restrict__ s, int n, int K, float* __restrict__ scal /*[0] weight_sum [1..3] var*/,
                                                          int* __restrict__ centres, int* __restrict__ n_centres) {
    __shared__ float s_v[kTbTile];
    const int job = blockIdx.x;
    const float it = (float)n, itm1 = (float)(n - 1);
    const float inv_it = 1.0f / it, inv_itm1 = 1.0f / itm1;     // div_s: multiply by the reciprocal (sutil/vec_math.h:483-487)
    float acc0 = 0.f, acc1 = 0.f, mean = 0.f;
    int nc = 0;
    for (int pass = 0; pass < 2; pass++) {
        for (int t0 = 0; t0 < n; t0 += kTbTile) {
            const int cnt = min(kTbTile, n - t0);
            for (int k = threadIdx.x; k < cnt; k += kTbBlock) {
                const spc_divide_weight& p = s[t0 + k];
                s_v[k] = job == 0 ? p.weight : (job == 1 ? p.position.x : (job == 2 ? p.position.y : p.position.z));
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                if (job == 0 && pass == 0) {
                    for (int k = 0; k < cnt; k++) acc0 += s_v[k];                       // weight_sum
                } else if (job == 0) {
                    const float quota = acc0 / K;                                         // weight_sum / K
                    for (int k = 0; k < cnt; k++) {
                        acc1 += s_v[k];
                        if (acc1 > quota) {
                            acc1 -= quota;
                            centres[nc++] = t0 + k;
                        }
                    }
                } else if (pass == 0) {
                    for (int k = 0; k < cnt; k++) mean = mean + s_v[k] * inv_it;
                } else {
                    for (int k = 0; k < cnt; k++) {
                        const float diff = mean - s_v[k];
                        acc1 = acc1 + (diff * diff) * inv_itm1;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        if (job == 0) {
            scal[0] = acc0;
            *n_centres = nc;
        } else {
            scal[job] = acc1;
        }
    }
}

// ---- nearest centre under divide_weight::d (classTree_common.h:82-90, DIR_JUDGE 0) ------------------------------------------------
__global__ void __launch_bounds__(256) k_tb_label(const spc_divide_weight* __restrict__ s, int n, const int* __restrict__ centres, int nc, int label_bias,
                                                  const float* __restrict__ scal, int* __restrict__ label, float* __restrict__ w_norm) {
    extern __shared__ float4 s_c[];          // per centre: {position, -} {normal, -} {dir, -}
    for (int c = threadIdx.x; c < nc; c += blockDim.x) {
        const spc_divide_weight& a = s[centres[c]];
        s_c[3 * c] = make_float4(a.position.x, a.position.y, a.position.z, 0.f);
        s_c[3 * c + 1] = make_float4(a.normal.x, a.normal.y, a.normal.z, 0.f);
        s_c[3 * c + 2] = make_float4(a.dir.x, a.dir.y, a.dir.z, 0.f);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float diversity2 = fmaxf(scal[1], fmaxf(scal[2], scal[3]));
    const float3 pp = ld3(s[i].position), pn = ld3(s[i].normal), pd = ld3(s[i].dir);
    float min_distance = FLT_MAX;
    int id = 0;
    for (int c = 0; c < nc; c++) {
        const float4 ap = s_c[3 * c], an = s_c[3 * c + 1], ad = s_c[3 * c + 2];
        const float dx = ap.x - pp.x, dy = ap.y - pp.y, dz = ap.z - pp.z;
        const float d_a = dx * dx + dy * dy + dz * dz;
        const float diff_direction = pd.x * ad.x + pd.y * ad.y + pd.z * ad.z;
        const float diff_normal = pn.x * an.x + pn.y * an.y + pn.z * an.z;
        const float d = d_a + diversity2 * ((1 - diff_normal) + (1 - diff_direction) * 0.0f);
        if (d < min_distance) {
            min_distance = d;
            id = c + label_bias;
        }
    }
    label[i] = id;
    w_norm[i] = s[i].weight / scal[0];       // para_initial: p.weight /= unnormal_weight (the same serial sum as weight_sum)
}

// bounding box of the positions; bbox_max starts at FLT_MIN (sic: the smallest positive float, classTree_host.h:99-100)
__global__ void k_tb_bbox(const spc_divide_weight* __restrict__ s, int n, float* __restrict__ bb /*[6] as ordered ints*/) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {FLT_MIN, FLT_MIN, FLT_MIN};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float p[3] = {s[i].position.x, s[i].position.y, s[i].position.z};
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], p[k]);
            hi[k] = fmaxf(hi[k], p[k]);
        }
    }
    // order-independent: min / max through the monotone int encoding of floats
    auto enc = [](float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; };
    for (int k = 0; k < 3; k++) {
        atomicMin(reinterpret_cast<int*>(bb) + k, enc(lo[k]));
        atomicMax(reinterpret_cast<int*>(bb) + 3 + k, enc(hi[k]));
    }
}
__global__ void k_tb_bbox_init(float* bb) {
    if (threadIdx.x < 3) reinterpret_cast<int*>(bb)[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) reinterpret_cast<int*>(bb)[threadIdx.x] = (int)0x80000000;
}
__device__ __forceinline__ float tb_dec(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ---- majority vote + weight of one run of samples (Builder::color, classTree_host.h:243-284) -------------------------------------
// `ord` lists the samples of the segment in their original relative order; `lord` the same samples stably sorted by label, `lpos`
// their positions in `ord`.  Called by a whole block; thread 0 does the serial work on tiles the block stages (see above) and
// holds the result.
struct Vote {
    float weight, cw;
    int   label;
    bool  pure;
};
struct VoteTile {
    float w[kTbTile];
    int   l[kTbTile];
    int   p[kTbTile];
    int   flag;
};
__device__ Vote tb_vote_block(VoteTile& sh, const int* __restrict__ ord, int b, int e, const int* __restrict__ lord, const int* __restrict__ lpos, int lb,
                              const int* __restrict__ label, const float* __restrict__ w, int inherit, bool sum_weight) {
    Vote v{0.f, 0.f, inherit, true};
    if (e <= b) return v;                                // (uniform over the block)
    const int first = label[ord[b]];
    float wsum = 0.f;
    bool pure = true;
    for (int t0 = b; t0 < e; t0 += kTbTile) {
        const int cnt = min(kTbTile, e - t0);
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            const int sid = ord[t0 + k];
            sh.w[k] = w[sid];
            sh.l[k] = label[sid];
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 0; k < cnt; k++) {
                if (sum_weight) wsum += sh.w[k];         // c.weight += s.weight, in the parent's sample order (:197-201)
                pure &= sh.l[k] == first;
            }
        __syncthreads();
    }
    if (threadIdx.x == 0) sh.flag = pure ? 1 : 0;
    __syncthreads();
    pure = sh.flag != 0;
    v.weight = wsum;
    v.pure = pure;
    v.label = first;
    if (pure) return v;                                  // cw = the node's weight: filled in by the caller
    // per-label sums in sample order; running maximum with "first to reach it wins"
    float best = 0.f, sum = 0.f;
    int best_label = first, best_pos = 0x7fffffff, cur = -1, reached = 0;
    const int le = lb + (e - b);
    auto close_run = [&]() {
        if (cur >= 0 && (sum > best || (sum == best && sum > 0.f && reached < best_pos))) {
            best = sum;
            best_label = cur;
            best_pos = reached;
        }
    };
    for (int t0 = lb; t0 < le; t0 += kTbTile) {
        const int cnt = min(kTbTile, le - t0);
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            const int sid = lord[t0 + k];
            sh.w[k] = w[sid];
            sh.l[k] = label[sid];
            sh.p[k] = lpos[t0 + k];
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 0; k < cnt; k++) {
                if (sh.l[k] != cur) {
                    close_run();
                    cur = sh.l[k];
                    sum = 0.f;
                    reached = sh.p[k];
                }
                const float ns = sum + sh.w[k];
                if (ns != sum) reached = sh.p[k];
                sum = ns;
            }
        __syncthreads();
    }
    close_run();
    v.cw = best;
    v.label = best_label;
    return v;
}

// root: one segment = all samples
__global__ void __launch_bounds__(kTbBlock) k_tb_root(const int* __restrict__ ord, int n, const int* __restrict__ lord, const int* __restrict__ lpos,
                                                      const int* __restrict__ label, const float* __restrict__ w, const float* __restrict__ bb, TbNode* __restrict__ nodes,
                                                      spc_tree_node* __restrict__ out, float* __restrict__ state) {
    __shared__ VoteTile sh;
    const Vote v = tb_vote_block(sh, ord, 0, n, lord, lpos, 0, label, w, 0, false);
    if (threadIdx.x != 0) return;
    TbNode r{};
    r.father = 0; r.depth = 0; r.pdepth = 0; r.ndepth = 0; r.slot = 0; r.begin = 0; r.end = n; r.cand = -1;
    r.weight = 1.f;                                      // v[0].weight = 1 (:360)
    r.cw = v.pure ? r.weight : v.cw;
    nodes[0] = r;
    spc_tree_node o{};
    o.leaf = 1;
    o.label = v.label;
    const float lo[3] = {tb_dec(reinterpret_cast<const int*>(bb)[0]), tb_dec(reinterpret_cast<const int*>(bb)[1]), tb_dec(reinterpret_cast<const int*>(bb)[2])};
    const float hi[3] = {tb_dec(reinterpret_cast<const int*>(bb)[3]), tb_dec(reinterpret_cast<const int*>(bb)[4]), tb_dec(reinterpret_cast<const int*>(bb)[5])};
    o.mid = spc_float3{(hi[0] + lo[0]) * 0.5f, (hi[1] + lo[1]) * 0.5f, (hi[2] + lo[2]) * 0.5f};   // div_s(bbox_max + bbox_min, 2)
    out[0] = o;
    state[0] = r.cw;      // c_w
}

// ---- per level -------------------------------------------------------------------------------------------------------------------
// which nodes of the level [lo, hi) are split candidates (flag), and their speculative split type and midpoint (Builder::split,
// :103-140: both depend on the ancestors only).  One thread per node.
__global__ void k_tb_candidates(TbNode* __restrict__ nodes, spc_tree_node* __restrict__ out, int lo, int hi, const float* __restrict__ bb, int* __restrict__ flags) {
    const int id = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= hi) return;
    TbNode& t = nodes[id];
    const bool cand = t.end > t.begin && t.cw < t.weight && t.depth < kMaxTreeDepth;
    flags[id - lo] = cand ? 1 : 0;
    if (!cand) return;
    const float blo[3] = {tb_dec(reinterpret_cast<const int*>(bb)[0]), tb_dec(reinterpret_cast<const int*>(bb)[1]), tb_dec(reinterpret_cast<const int*>(bb)[2])};
    const float bhi[3] = {tb_dec(reinterpret_cast<const int*>(bb)[3]), tb_dec(reinterpret_cast<const int*>(bb)[4]), tb_dec(reinterpret_cast<const int*>(bb)[5])};
    const int split_type = (t.depth % 2 == 0 || t.ndepth > 3) ? 0 : 1;
    float inch[3];
    if (split_type == 0) {
        for (int k = 0; k < 3; k++) {
            float b = bhi[k] - blo[k];
            for (int j = 0; j < t.pdepth + 1; j++) b = b * 0.5f;     // block_size[position_depth + 1]
            inch[k] = b;
        }
    } else {
        float b = 2.0f;
        for (int j = 0; j < t.ndepth + 1; j++) b = b * 0.5f;         // direction_block_size[normal_depth + 1]
        inch[0] = inch[1] = inch[2] = b;
    }
    float mid[3];
    if (t.ndepth == 0 && split_type == 1) {
        mid[0] = mid[1] = mid[2] = 0.f;
    } else if (t.pdepth == 0) {
        mid[0] = out[id].mid.x; mid[1] = out[id].mid.y; mid[2] = out[id].mid.z;
    } else {
        int L_id = id, t_id = t.father;
        while (t_id != 0 && out[t_id].type != split_type) {
            L_id = t_id;
            t_id = nodes[t_id].father;
        }
        const int c = nodes[L_id].slot;
        mid[0] = out[t_id].mid.x + ((c >> 0) % 2 == 0 ? -inch[0] : inch[0]);
        mid[1] = out[t_id].mid.y + ((c >> 1) % 2 == 0 ? -inch[1] : inch[1]);
        mid[2] = out[t_id].mid.z + ((c >> 2) % 2 == 0 ? -inch[2] : inch[2]);
    }
    // kept in the output record; becomes final only if the serial pass accepts the split (the leaf flag stays 1 until then)
    out[id].mid = spc_float3{mid[0], mid[1], mid[2]};
    out[id].type = split_type;
}
// flags -> ranks: cand = rank among the level's candidates (or -1), cand_node[rank] = node id
__global__ void k_tb_rank(TbNode* __restrict__ nodes, int lo, int hi, const int* __restrict__ flags, const int* __restrict__ ranks, int* __restrict__ cand_node) {
    const int id = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= hi) return;
    const int f = flags[id - lo], r = ranks[id - lo];
    nodes[id].cand = f ? r : -1;
    if (f) cand_node[r] = id;
}
// node id of every position of the level's sample order (positions of node i: [begin, end))
__global__ void k_tb_pos_node(const TbNode* __restrict__ nodes, int lo, int hi, int* __restrict__ pos_node) {
    const int id = lo + blockIdx.x;
    if (id >= hi) return;
    for (int k = nodes[id].begin + threadIdx.x; k < nodes[id].end; k += blockDim.x) pos_node[k] = id;
}
// key of every sample of the level: (candidate rank, child slot), or all-ones for samples whose node is not split
__global__ void k_tb_slot_keys(const int* __restrict__ ord, int m, const int* __restrict__ pos_node, const TbNode* __restrict__ nodes, const spc_tree_node* __restrict__ out,
                               const spc_divide_weight* __restrict__ s, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int sid = ord[k];
    vals[k] = sid;
    const int id = pos_node[k];                      // -1: the sample dropped out at an earlier level
    const int cand = id >= 0 ? nodes[id].cand : -1;
    if (cand < 0) {
        keys[k] = ~0ull;
        return;
    }
    const spc_tree_node& nd = out[id];
    const spc_float3 q = nd.type == 0 ? s[sid].position : (nd.type == 1 ? s[sid].normal : s[sid].dir);   // tree_node::operator(), classTree_common.h:20-36
    int ind = 0;
    ind += q.x > nd.mid.x ? 1 : 0;
    ind += q.y > nd.mid.y ? 2 : 0;
    ind += q.z > nd.mid.z ? 4 : 0;
    keys[k] = ((unsigned long long)cand << 3) | (unsigned long long)ind;
}
// label keys of the (already child-sorted) samples: (child segment, label); values = position in the child-sorted order
__global__ void k_tb_label_keys(const unsigned long long* __restrict__ slot_keys, const int* __restrict__ ord, int m, const int* __restrict__ label, int label_bits,
                                unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    keys[k] = (slot_keys[k] << label_bits) | (unsigned long long)label[ord[k]];
    vals[k] = k;
}
__device__ __forceinline__ int tb_lower_bound(const unsigned long long* __restrict__ a, int n, unsigned long long v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
// one BLOCK per speculative child (candidate rank * 8 + slot): its sample range, weight, majority label
struct TbChild {
    int   begin, end;
    float weight, cw;
    int   label;
};
__global__ void __launch_bounds__(kTbBlock) k_tb_children(const unsigned long long* __restrict__ slot_keys, const int* __restrict__ ord, int m,
                                                          const unsigned long long* __restrict__ lkeys, const int* __restrict__ lpos, int label_bits,
                                                          const int* __restrict__ label, const float* __restrict__ w, int n_children, int* __restrict__ lord_scratch,
                                                          TbChild* __restrict__ children) {
    __shared__ VoteTile sh;
    const int c = blockIdx.x;
    if (c >= n_children) return;
    const int b = tb_lower_bound(slot_keys, m, (unsigned long long)c), e = tb_lower_bound(slot_keys, m, (unsigned long long)c + 1);
    const int lb = tb_lower_bound(lkeys, m, (unsigned long long)c << label_bits);
    // lord = sample ids in label order: ord[lpos[k]]
    for (int k = lb + threadIdx.x; k < lb + (e - b); k += blockDim.x) lord_scratch[k] = ord[lpos[k]];
    __syncthreads();
    const Vote v = tb_vote_block(sh, ord, b, e, lord_scratch, lpos, lb, label, w, -1, true);
    if (threadIdx.x != 0) return;
    TbChild ch;
    ch.begin = b; ch.end = e; ch.weight = v.weight; ch.label = v.label;
    ch.cw = (e > b) ? (v.pure ? v.weight : v.cw) : 0.f;
    children[c] = ch;
}
// per candidate: what accepting its split would add to c_w = the sum of its children's majority weights, in child order (:203-206)
__global__ void k_tb_child_sums(const int* __restrict__ cand_node, const TbNode* __restrict__ nodes, const TbChild* __restrict__ children, int n_cand,
                                float* __restrict__ cand_cw, float* __restrict__ cand_sum) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_cand) return;
    float sum = 0.0f;
    for (int i = 0; i < 8; i++) sum += children[r * 8 + i].cw;
    cand_sum[r] = sum;
    cand_cw[r] = nodes[cand_node[r]].cw;
}
// the reference's loop over the level (classTree_host.h:364-371), on the candidates in node order: a split counts while 0.99 > c_w;
// c_w -= correct_weight; c_w += split().  Serial by nature (one thread adds, the block stages).
__global__ void __launch_bounds__(kTbBlock) k_tb_accept(const float* __restrict__ cand_cw, const float* __restrict__ cand_sum, int n_cand, float* __restrict__ state,
                                                        int* __restrict__ accepted) {
    __shared__ float s_cw[kTbTile], s_sum[kTbTile];
    __shared__ int s_acc[kTbTile];
    float c_w = state[0];
    for (int t0 = 0; t0 < n_cand; t0 += kTbTile) {
        const int cnt = min(kTbTile, n_cand - t0);
        for (int k = threadIdx.x; k < cnt; k += kTbBlock) {
            s_cw[k] = cand_cw[t0 + k];
            s_sum[k] = cand_sum[t0 + k];
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 0; k < cnt; k++) {
                const bool ok = kAccuracy > c_w;
                s_acc[k] = ok ? 1 : 0;
                if (ok) {
                    c_w -= s_cw[k];
                    c_w += s_sum[k];
                }
            }
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += kTbBlock) accepted[t0 + k] = s_acc[k];
        __syncthreads();
    }
    if (threadIdx.x == 0) state[0] = c_w;
}
// accepted splits become inner nodes with 8 children appended in candidate order; everything else on the level stays a leaf
__global__ void k_tb_make_children(TbNode* __restrict__ nodes, spc_tree_node* __restrict__ out, int lo, int hi, const TbChild* __restrict__ children,
                                   const int* __restrict__ accepted, const int* __restrict__ acc_rank, int first_new) {
    const int id = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= hi) return;
    const TbNode t = nodes[id];
    if (t.cand < 0 || !accepted[t.cand]) {
        // a leaf keeps the zero mid / type of a fresh node (Node()), except the root's mid (:361)
        if (id != 0) out[id].mid = spc_float3{0.f, 0.f, 0.f};
        out[id].type = 0;
        return;
    }
    const int back = first_new + 8 * acc_rank[t.cand];
    out[id].leaf = 0;
    const int type = out[id].type;
    for (int i = 0; i < 8; i++) {
        const TbChild ch = children[t.cand * 8 + i];
        out[id].child[i] = back + i;
        TbNode c{};
        c.father = id; c.depth = t.depth + 1; c.slot = i; c.cand = -1;
        c.pdepth = t.pdepth + (type == 0);
        c.ndepth = t.ndepth + (type == 1);
        c.begin = ch.begin; c.end = ch.end;
        c.weight = ch.weight;
        c.cw = ch.cw;
        nodes[back + i] = c;
        spc_tree_node o{};
        o.leaf = 1;
        o.label = ch.end > ch.begin ? ch.label : out[id].label;     // an empty child keeps its father's label (:177)
        out[back + i] = o;
    }
}
__global__ void k_tb_max_label(const spc_tree_node* __restrict__ out, int n, int* __restrict__ max_label) {
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, out[i].label);
    atomicMax(max_label, m);
}

struct SortTmp {
    DevBuf<uint8_t> buf;
    void pairs(cudaStream_t st, const unsigned long long* kin, unsigned long long* kout, const int* vin, int* vout, int n, int bits) {
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, bits, st);
        buf.alloc(bytes);
        SPC_CUDA(cub::DeviceRadixSort::SortPairs(buf.p, bytes, kin, kout, vin, vout, n, 0, bits, st));
    }
};

int bits_of(unsigned long long v) {
    int b = 1;
    while (b < 64 && (v >> b)) b++;
    return b;
}

}  // namespace

// Builds the tree of `n` device-resident samples; returns the node count; `out` receives the nodes (device).
int tree_build_device(Context& c, const spc_divide_weight* samples, int n, int K, int label_bias, DevBuf<spc_tree_node>& out, int* max_label_host) {
    NvtxRange range("spc: tree build (device)");
    SPC_REQUIRE(samples && n >= 2 && K >= 1, SPC_ERR_INVALID, "tree build: bad arguments (n=%d, K=%d)", n, K);
    cudaStream_t st = c.stream;
    DevBuf<float> scal, w, bb, state, cand_cw, cand_sum;
    DevBuf<int> centres, counters, label, ord, ord2, pos_node, lvals, lvals2, lord, flags, ranks, cand_node, accepted, acc_rank;
    DevBuf<unsigned long long> keys, keys2, lkeys, lkeys2;
    DevBuf<TbNode> nodes;
    DevBuf<TbChild> children;
    DevBuf<uint8_t> scan_tmp;
    SortTmp sorter;
    scal.alloc(4); bb.alloc(6); state.alloc(2); centres.alloc(K + 1); counters.alloc(8);
    w.alloc(n); label.alloc(n); ord.alloc(n); ord2.alloc(n); pos_node.alloc(n); lvals.alloc(n); lvals2.alloc(n); lord.alloc(n);
    keys.alloc(n); keys2.alloc(n); lkeys.alloc(n); lkeys2.alloc(n);
    SPC_CUDA(cudaMemsetAsync(counters.p, 0, 8 * sizeof(int), st));
    int* d_n_centres = counters.p, *d_max_label = counters.p + 4;
    auto exclusive_sum = [&](const int* in, int* outp, int count) {
        size_t bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, outp, count, st);
        scan_tmp.alloc(bytes);
        SPC_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp.p, bytes, in, outp, count, st));
    };
    auto last_plus = [&](const int* flags_p, const int* ranks_p, int count) {   // total of an exclusive scan = last rank + last flag
        int h[2] = {0, 0};
        SPC_CUDA(cudaMemcpyAsync(&h[0], flags_p + count - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        SPC_CUDA(cudaMemcpyAsync(&h[1], ranks_p + count - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
        SPC_CUDA(cudaStreamSynchronize(st));
        return h[0] + h[1];
    };

    k_tb_prologue<<<4, kTbBlock, 0, st>>>(samples, n, K, scal.p, centres.p, d_n_centres);
    k_tb_bbox_init<<<1, 32, 0, st>>>(bb.p);
    k_tb_bbox<<<c.sm_count, 256, 0, st>>>(samples, n, bb.p);
    int nc = 0;
    SPC_CUDA(cudaMemcpyAsync(&nc, d_n_centres, sizeof(int), cudaMemcpyDeviceToHost, st));
    SPC_CUDA(cudaStreamSynchronize(st));
    SPC_REQUIRE(nc >= 1 && nc <= K + 1, SPC_ERR_INVALID, "tree build: %d centres for K = %d", nc, K);
    const size_t smem = (size_t)nc * 3 * sizeof(float4);
    SPC_REQUIRE(smem <= 200 * 1024, SPC_ERR_CAPACITY, "tree build: %d centres do not fit shared memory", nc);
    SPC_CUDA(cudaFuncSetAttribute(k_tb_label, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_tb_label<<<(n + 255) / 256, 256, smem, st>>>(samples, n, centres.p, nc, label_bias, scal.p, label.p, w.p);
    c.launches += 4;

    // root: identity order; label-sorted view for the majority vote
    const int label_bits = bits_of((unsigned long long)(nc + label_bias));
    {
        std::vector<int> iota((size_t)n);
        for (int i = 0; i < n; i++) iota[i] = i;
        SPC_CUDA(cudaMemcpyAsync(ord.p, iota.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
        SPC_CUDA(cudaStreamSynchronize(st));
    }
    SPC_CUDA(cudaMemsetAsync(keys.p, 0, (size_t)n * sizeof(unsigned long long), st));   // one segment: slot key 0 everywhere
    k_tb_label_keys<<<(n + 255) / 256, 256, 0, st>>>(keys.p, ord.p, n, label.p, label_bits, lkeys.p, lvals.p);
    sorter.pairs(st, lkeys.p, lkeys2.p, lvals.p, lvals2.p, n, label_bits + 1);
    int capacity = 1 + 8 * 4096;
    nodes.alloc(capacity);
    out.alloc(capacity);
    // for the identity order, position == sample id: the label-sorted positions serve as both `lord` and `lpos`
    k_tb_root<<<1, kTbBlock, 0, st>>>(ord.p, n, lvals2.p, lvals2.p, label.p, w.p, bb.p, nodes.p, out.p, state.p);
    c.launches += 3;

    int lo = 0, hi = 1;             // current level = nodes [lo, hi), whose samples are positions of `ord` (m of them, some dropped)
    const int m = n;
    for (int depth = 0; depth < kMaxTreeDepth && hi > lo; depth++) {
        const int n_level = hi - lo;
        flags.alloc(n_level); ranks.alloc(n_level);
        k_tb_candidates<<<(n_level + 127) / 128, 128, 0, st>>>(nodes.p, out.p, lo, hi, bb.p, flags.p);
        exclusive_sum(flags.p, ranks.p, n_level);
        const int n_cand = last_plus(flags.p, ranks.p, n_level);
        c.launches += 2;
        if (n_cand == 0) break;
        cand_node.alloc(n_cand); cand_cw.alloc(n_cand); cand_sum.alloc(n_cand); accepted.alloc(n_cand); acc_rank.alloc(n_cand);
        k_tb_rank<<<(n_level + 127) / 128, 128, 0, st>>>(nodes.p, lo, hi, flags.p, ranks.p, cand_node.p);
        // room for every speculative child
        if (hi + 8 * n_cand > capacity) {
            const int new_cap = std::max(capacity * 2, hi + 8 * n_cand);
            DevBuf<TbNode> nn;
            DevBuf<spc_tree_node> no;
            nn.alloc(new_cap); no.alloc(new_cap);
            SPC_CUDA(cudaMemcpyAsync(nn.p, nodes.p, (size_t)hi * sizeof(TbNode), cudaMemcpyDeviceToDevice, st));
            SPC_CUDA(cudaMemcpyAsync(no.p, out.p, (size_t)hi * sizeof(spc_tree_node), cudaMemcpyDeviceToDevice, st));
            SPC_CUDA(cudaStreamSynchronize(st));
            std::swap(nodes.p, nn.p); std::swap(nodes.n, nn.n);
            std::swap(out.p, no.p); std::swap(out.n, no.n);
            capacity = new_cap;
        }
        children.alloc((size_t)n_cand * 8);
        SPC_CUDA(cudaMemsetAsync(pos_node.p, 0xff, (size_t)m * sizeof(int), st));
        k_tb_pos_node<<<n_level, 128, 0, st>>>(nodes.p, lo, hi, pos_node.p);
        k_tb_slot_keys<<<(m + 255) / 256, 256, 0, st>>>(ord.p, m, pos_node.p, nodes.p, out.p, samples, keys.p, lvals.p);
        // stable sort by (candidate, slot): children's samples become contiguous and keep their relative order; samples of nodes that
        // are not split (key all-ones: only the low bits are sorted, all ones there too) go to the end and drop out
        const int slot_bits = bits_of(((unsigned long long)n_cand << 3)) + 1;
        sorter.pairs(st, keys.p, keys2.p, lvals.p, ord2.p, m, slot_bits);
        k_tb_label_keys<<<(m + 255) / 256, 256, 0, st>>>(keys2.p, ord2.p, m, label.p, label_bits, lkeys.p, lvals.p);
        sorter.pairs(st, lkeys.p, lkeys2.p, lvals.p, lvals2.p, m, std::min(64, slot_bits + label_bits));
        k_tb_children<<<n_cand * 8, kTbBlock, 0, st>>>(keys2.p, ord2.p, m, lkeys2.p, lvals2.p, label_bits, label.p, w.p, n_cand * 8, lord.p, children.p);
        k_tb_child_sums<<<(n_cand + 127) / 128, 128, 0, st>>>(cand_node.p, nodes.p, children.p, n_cand, cand_cw.p, cand_sum.p);
        k_tb_accept<<<1, kTbBlock, 0, st>>>(cand_cw.p, cand_sum.p, n_cand, state.p, accepted.p);
        exclusive_sum(accepted.p, acc_rank.p, n_cand);
        const int n_acc = last_plus(accepted.p, acc_rank.p, n_cand);
        k_tb_make_children<<<(n_level + 127) / 128, 128, 0, st>>>(nodes.p, out.p, lo, hi, children.p, accepted.p, acc_rank.p, hi);
        c.launches += 11;
        SPC_CUDA(cudaGetLastError());
        std::swap(ord.p, ord2.p);
        std::swap(ord.n, ord2.n);
        lo = hi;
        hi = hi + 8 * n_acc;
        // (the dropped samples stay at the end of the order with no node: k_tb_slot_keys gives them the all-ones key again)
    }
    const int n_nodes = hi;
    if (max_label_host) {
        k_tb_max_label<<<c.sm_count, 256, 0, st>>>(out.p, n_nodes, d_max_label);
        SPC_CUDA(cudaMemcpyAsync(max_label_host, d_max_label, sizeof(int), cudaMemcpyDeviceToHost, st));
        c.launches++;
    }
    SPC_CUDA(cudaStreamSynchronize(st));
    SPC_CUDA(cudaGetLastError());
    return n_nodes;
}

}  // namespace spc
