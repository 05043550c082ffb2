// bvh_build.cu -- GPU builder of the compressed 8-wide BVH.  Replaces optixAccelBuild for the GAS
// per mesh + single-level IAS of the reference (sutil/Scene.cpp:943-1257, 1260-1338): all instance
// transforms there are the identity (scene_shift.cpp:241,322), so the scene is flattened to one level.
//
// Pipeline (all on the device, one stream):
//   1. per-triangle bounds + scene bounds        (k_prim_bounds)
//   2. 63-bit Morton codes of box centres        (k_morton)  -> cub radix sort
//   3. binary radix tree, Karras 2012            (k_radix_tree)
//   4. bottom-up refit fused with the SAH-optimal wide-collapse DP of Ylitie et al. 2017
//      (cost table c(n,1..7) per binary node)    (k_fit_dp)
//   5. top-down emission of 80-byte nodes + 48-byte triangles, one queue per level (k_emit)
#include <cub/cub.cuh>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include "traverse.cuh"

namespace spc {

namespace {

__device__ float g_cost_prim = 0.3f;   // SAH cost of one triangle test relative to one node step (SPC_BVH_COST_PRIM overrides)
#define kCostPrim g_cost_prim
constexpr float kCostNode = 1.0f;
constexpr int   kMaxLeafTris = 3;

__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

// 1 ------------------------------------------------------------------------------------------
__global__ void k_prim_bounds(const float4* __restrict__ tri_pos, uint32_t n, float4* __restrict__ plo,
                              float4* __restrict__ phi, uint32_t* __restrict__ scene /*6 ordered uints*/) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        const float4 a = tri_pos[3 * (size_t)i], b = tri_pos[3 * (size_t)i + 1], c = tri_pos[3 * (size_t)i + 2];
        lo[0] = fminf(a.x, fminf(b.x, c.x)); hi[0] = fmaxf(a.x, fmaxf(b.x, c.x));
        lo[1] = fminf(a.y, fminf(b.y, c.y)); hi[1] = fmaxf(a.y, fmaxf(b.y, c.y));
        lo[2] = fminf(a.z, fminf(b.z, c.z)); hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
        plo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        phi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    for (int k = 0; k < 3; k++) {
        float l = lo[k], h = hi[k];
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(scene + k, f2ord(l));
            atomicMax(scene + 3 + k, f2ord(h));
        }
    }
}

// 2 ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread21(uint32_t v) {
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__global__ void k_morton(const float4* __restrict__ plo, const float4* __restrict__ phi, uint32_t n,
                         float3 slo, float3 sinv, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 l = plo[i], h = phi[i];
    const float cx = (0.5f * (l.x + h.x) - slo.x) * sinv.x;
    const float cy = (0.5f * (l.y + h.y) - slo.y) * sinv.y;
    const float cz = (0.5f * (l.z + h.z) - slo.z) * sinv.z;
    const uint32_t qx = (uint32_t)fminf(fmaxf(cx * 2097152.f, 0.f), 2097151.f);
    const uint32_t qy = (uint32_t)fminf(fmaxf(cy * 2097152.f, 0.f), 2097151.f);
    const uint32_t qz = (uint32_t)fminf(fmaxf(cz * 2097152.f, 0.f), 2097151.f);
    keys[i] = (spread21(qx) << 2) | (spread21(qy) << 1) | spread21(qz);
    vals[i] = i;
}

// 3 ------------------------------------------------------------------------------------------
// node ids: internal 0..n-2 (root 0), leaf k -> n-1+k.
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t x = keys[i] ^ keys[j];
    return x ? __clzll((long long)x) : 64 + __clz(i ^ j);
}
__global__ void k_radix_tree(const uint64_t* __restrict__ keys, int n, int* __restrict__ left,
                             int* __restrict__ right, int* __restrict__ parent, int* __restrict__ first,
                             int* __restrict__ last) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    left[i] = lc;
    right[i] = rc;
    parent[lc] = i;
    parent[rc] = i;
    first[i] = lo;
    last[i] = hi;
    if (i == 0) parent[0] = -1;
}

// 3b -----------------------------------------------------------------------------------------
// Binary topology by parallel locally-ordered clustering (PLOC, Meister & Bittner 2018) instead of the Morton radix
// tree: clusters stay in Morton order; every round each cluster looks kPlocRadius neighbours to either side for the one
// whose merged box has the smallest area, mutual nearest neighbours merge, the array is compacted (order kept).
// Deterministic: ties go to the lower index and node ids come from a prefix sum, not from an atomic counter.
// The radix tree only groups by Morton prefix; merging by surface area gives a tree whose SAH cost is markedly lower
// (fewer node visits per ray), and the hit results cannot change (intersection contract, traverse.cuh).
// Output in the form steps 4-5 expect: internal ids 0..n-2 with root 0, leaf k -> n-1+k where k is the position
// in a depth-first leaf order (so every subtree owns a contiguous [first,last] range of `sorted_prim`).
constexpr int kPlocRadius = 16;

__global__ void k_ploc_init(int n, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ plo,
                            const float4* __restrict__ phi, float pad, float4* nb_lo, float4* nb_hi, int* cid, int* cnt) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t p = sorted_prim[k];
    float4 l = plo[p], h = phi[p];
    l.x -= pad; l.y -= pad; l.z -= pad;
    h.x += pad; h.y += pad; h.z += pad;
    nb_lo[n - 1 + k] = l;
    nb_hi[n - 1 + k] = h;
    cid[k] = n - 1 + k;
    cnt[n - 1 + k] = 1;
}

__global__ void k_ploc_nn(int m, const int* __restrict__ cid, const float4* __restrict__ nb_lo,
                          const float4* __restrict__ nb_hi, int* __restrict__ nn) {
    extern __shared__ float4 s_box[];   // [blockDim + 2R] lo, then hi
    const int R = kPlocRadius;
    const int W = blockDim.x + 2 * R;
    float4* s_lo = s_box;
    float4* s_hi = s_box + W;
    const int base = blockIdx.x * blockDim.x - R;
    for (int t = threadIdx.x; t < W; t += blockDim.x) {
        const int g = base + t;
        if (g >= 0 && g < m) {
            const int id = cid[g];
            s_lo[t] = nb_lo[id];
            s_hi[t] = nb_hi[id];
        }
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int li = threadIdx.x + R;
    const float4 a_lo = s_lo[li], a_hi = s_hi[li];
    float best = FLT_MAX;
    int bj = -1;
    for (int d = -R; d <= R; d++) {
        const int j = i + d;
        if (d == 0 || j < 0 || j >= m) continue;
        const float4 b_lo = s_lo[li + d], b_hi = s_hi[li + d];
        const float dx = fmaxf(a_hi.x, b_hi.x) - fminf(a_lo.x, b_lo.x);
        const float dy = fmaxf(a_hi.y, b_hi.y) - fminf(a_lo.y, b_lo.y);
        const float dz = fmaxf(a_hi.z, b_hi.z) - fminf(a_lo.z, b_lo.z);
        const float area = dx * dy + dy * dz + dz * dx;
        if (area < best) { best = area; bj = j; }   // ascending j: ties keep the lower index
    }
    nn[i] = bj;
}

// flag word: low 32 bits = 1 when position i survives the round, high 32 bits = 1 when i is the lower half of a merging pair
__global__ void k_ploc_flags(int m, const int* __restrict__ nn, unsigned long long* __restrict__ fl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int j = nn[i];
    const bool mutual = j >= 0 && nn[j] == i;
    unsigned long long f = 1ull;
    if (mutual) f = (i < j) ? (1ull | (1ull << 32)) : 0ull;
    fl[i] = f;
}

__global__ void k_ploc_apply(int m, int n, int merged_before, const int* __restrict__ cid, const int* __restrict__ nn,
                             const unsigned long long* __restrict__ fl, const unsigned long long* __restrict__ sc,
                             int* __restrict__ cid_out, int* left, int* right, int* parent, int* cnt, float4* nb_lo, float4* nb_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const unsigned long long f = fl[i];
    if (!(f & 1ull)) return;
    const unsigned long long s = sc[i];
    const int pos = (int)(s & 0xffffffffull);
    int id = cid[i];
    if (f >> 32) {
        const int rank = (int)(s >> 32);
        const int a = id, b = cid[nn[i]];
        id = n - 2 - (merged_before + rank);
        left[id] = a;
        right[id] = b;
        parent[a] = id;
        parent[b] = id;
        cnt[id] = cnt[a] + cnt[b];
        const float4 al = nb_lo[a], ah = nb_hi[a], bl = nb_lo[b], bh = nb_hi[b];
        nb_lo[id] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f);
        nb_hi[id] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
    }
    cid_out[pos] = id;
}

// depth-first position of every node's first leaf: sum of the left siblings' leaf counts on the way to the root
__global__ void k_ploc_order(int n, const int* __restrict__ left, const int* __restrict__ parent, const int* __restrict__ cnt,
                             const uint32_t* __restrict__ sorted_prim, uint32_t* __restrict__ sorted_prim_out,
                             int* __restrict__ first, int* __restrict__ last, int* __restrict__ leaf_pos) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * n - 1) return;
    int pos = 0;
    int c = id;
    int p = parent[c];
    while (p >= 0) {
        const int l = left[p];
        if (l != c) pos += cnt[l];
        c = p;
        p = parent[c];
    }
    if (id >= n - 1) {
        leaf_pos[id - (n - 1)] = pos;
        sorted_prim_out[pos] = sorted_prim[id - (n - 1)];
    } else {
        first[id] = pos;
        last[id] = pos + cnt[id] - 1;
    }
}

// renumber the leaves by their depth-first position
__global__ void k_ploc_relink(int n, const int* __restrict__ leaf_pos, int* left, int* right, const int* __restrict__ parent,
                              int* __restrict__ parent_out) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= 2 * n - 1) return;
    if (id >= n - 1) {
        parent_out[n - 1 + leaf_pos[id - (n - 1)]] = parent[id];
    } else {
        parent_out[id] = parent[id];
        const int l = left[id], r = right[id];
        if (l >= n - 1) left[id] = n - 1 + leaf_pos[l - (n - 1)];
        if (r >= n - 1) right[id] = n - 1 + leaf_pos[r - (n - 1)];
    }
}

// 4 ------------------------------------------------------------------------------------------
__device__ __forceinline__ float half_area(const float4 lo, const float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}
// dec[m*8+0] = k of distribute(m,8); dec[m*8+i-1], i=2..7 = k of distribute(m,i) or 0 ("use i-1");
// dec[m*8+7] = 1 when c(m,1) is the leaf alternative.
__global__ void k_fit_dp(int n, const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ plo,
                         const float4* __restrict__ phi, float pad, const int* __restrict__ left,
                         const int* __restrict__ right, const int* __restrict__ parent,
                         const int* __restrict__ first, const int* __restrict__ last, float4* nb_lo,
                         float4* nb_hi, float* cost, uint8_t* dec, int* flags) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    {
        const uint32_t p = sorted_prim[k];
        float4 l = plo[p], h = phi[p];
        l.x -= pad; l.y -= pad; l.z -= pad;
        h.x += pad; h.y += pad; h.z += pad;
        nb_lo[n - 1 + k] = l;
        nb_hi[n - 1 + k] = h;
    }
    __threadfence();
    int cur = (n > 1) ? parent[n - 1 + k] : -1;
    while (cur >= 0) {
        if (atomicAdd(flags + cur, 1) == 0) return;
        const int lc = left[cur], rc = right[cur];
        const float4 ll = __ldcg(nb_lo + lc), lh = __ldcg(nb_hi + lc);
        const float4 rl = __ldcg(nb_lo + rc), rh = __ldcg(nb_hi + rc);
        const float4 lo = make_float4(fminf(ll.x, rl.x), fminf(ll.y, rl.y), fminf(ll.z, rl.z), 0.f);
        const float4 hi = make_float4(fmaxf(lh.x, rh.x), fmaxf(lh.y, rh.y), fmaxf(lh.z, rh.z), 0.f);
        nb_lo[cur] = lo;
        nb_hi[cur] = hi;
        float cL[8], cR[8];
        if (lc >= n - 1) {
            const float a = half_area(ll, lh) * kCostPrim;
            for (int i = 1; i <= 7; i++) cL[i] = a;
        } else
            for (int i = 1; i <= 7; i++) cL[i] = __ldcg(cost + (size_t)lc * 7 + i - 1);
        if (rc >= n - 1) {
            const float a = half_area(rl, rh) * kCostPrim;
            for (int i = 1; i <= 7; i++) cR[i] = a;
        } else
            for (int i = 1; i <= 7; i++) cR[i] = __ldcg(cost + (size_t)rc * 7 + i - 1);
        float dist[9];
        int   dk[9];
        for (int j = 2; j <= 8; j++) {
            float best = FLT_MAX;
            int   bk = 1;
            for (int kk = 1; kk < j; kk++) {
                if (kk > 7 || j - kk > 7) continue;
                const float c = cL[kk] + cR[j - kk];
                if (c < best) { best = c; bk = kk; }
            }
            dist[j] = best;
            dk[j] = bk;
        }
        const float A = half_area(lo, hi);
        const int   P = last[cur] - first[cur] + 1;
        const float c_leaf = (P <= kMaxLeafTris) ? A * P * kCostPrim : FLT_MAX;
        const float c_int = dist[8] + A * kCostNode;
        float c[8];
        uint8_t* dd = dec + (size_t)cur * 8;
        c[1] = fminf(c_leaf, c_int);
        dd[7] = (c_leaf <= c_int) ? 1 : 0;
        dd[0] = (uint8_t)dk[8];
        for (int i = 2; i <= 7; i++) {
            if (dist[i] < c[i - 1]) { c[i] = dist[i]; dd[i - 1] = (uint8_t)dk[i]; }
            else { c[i] = c[i - 1]; dd[i - 1] = 0; }
        }
        for (int i = 1; i <= 7; i++) cost[(size_t)cur * 7 + i - 1] = c[i];
        __threadfence();
        cur = parent[cur];
    }
}

// 5 ------------------------------------------------------------------------------------------
struct EmitItem {
    int node;   // binary node id
    int wide;   // index of the 8-wide node to write
    int depth;
};
struct EmitCounters {
    uint32_t n_wide;     // allocated 8-wide nodes
    uint32_t n_tris;     // allocated output triangles
    uint32_t n_next;     // items pushed for the next level
    uint32_t max_depth;
    float    sah;        // sum of (area * cost) -- divided by root area on the host
};

__global__ void k_emit(int n, const EmitItem* __restrict__ items, int n_items, EmitItem* __restrict__ next,
                       EmitCounters* ctr, const int* __restrict__ left, const int* __restrict__ right,
                       const int* __restrict__ first, const int* __restrict__ last,
                       const uint8_t* __restrict__ dec, const float4* __restrict__ nb_lo,
                       const float4* __restrict__ nb_hi, const uint32_t* __restrict__ sorted_prim,
                       const float4* __restrict__ tri_pos, float4* __restrict__ out_nodes,
                       float4* __restrict__ out_tris) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_items) return;
    const EmitItem it = items[w];
    const int nl = n - 1;  // first leaf id

    // ---- collect up to 8 children by following the DP decisions
    int c_node[8], c_first[8], c_cnt[8];
    bool c_inner[8];
    int nc = 0;
    {
        int st_node[16], st_budget[16];
        bool st_force[16];
        int sp = 0;
        if (it.node >= nl) {
            c_node[0] = it.node; c_first[0] = it.node - nl; c_cnt[0] = 1; c_inner[0] = false; nc = 1;
        } else if (last[it.node] - first[it.node] + 1 <= kMaxLeafTris && dec[(size_t)it.node * 8 + 7]) {
            // root that the DP would rather keep as one leaf
            c_node[0] = it.node; c_first[0] = first[it.node]; c_cnt[0] = last[it.node] - first[it.node] + 1;
            c_inner[0] = false; nc = 1;
        } else {
            st_node[0] = it.node; st_budget[0] = 8; st_force[0] = true; sp = 1;
        }
        while (sp > 0) {
            sp--;
            const int m = st_node[sp];
            int b = st_budget[sp];
            const bool force = st_force[sp];
            if (m >= nl) {
                c_node[nc] = m; c_first[nc] = m - nl; c_cnt[nc] = 1; c_inner[nc] = false; nc++;
                continue;
            }
            const uint8_t* dd = dec + (size_t)m * 8;
            int k;
            if (force) k = dd[0];
            else {
                while (b > 1 && dd[b - 1] == 0) b--;
                if (b == 1) {
                    c_node[nc] = m;
                    if (dd[7]) { c_first[nc] = first[m]; c_cnt[nc] = last[m] - first[m] + 1; c_inner[nc] = false; }
                    else { c_first[nc] = 0; c_cnt[nc] = 0; c_inner[nc] = true; }
                    nc++;
                    continue;
                }
                k = dd[b - 1];
            }
            st_node[sp] = right[m]; st_budget[sp] = b - k; st_force[sp] = false; sp++;
            st_node[sp] = left[m];  st_budget[sp] = k;     st_force[sp] = false; sp++;
        }
    }

    // ---- node box and quantisation frame
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float clo[8][3], chi[8][3];
    for (int c = 0; c < nc; c++) {
        const float4 l = nb_lo[c_node[c]], h = nb_hi[c_node[c]];
        clo[c][0] = l.x; clo[c][1] = l.y; clo[c][2] = l.z;
        chi[c][0] = h.x; chi[c][1] = h.y; chi[c][2] = h.z;
        for (int a = 0; a < 3; a++) { lo[a] = fminf(lo[a], clo[c][a]); hi[a] = fmaxf(hi[a], chi[c][a]); }
    }
    int   ebias[3];
    float scale[3];
    for (int a = 0; a < 3; a++) {
        const float ext = fmaxf(hi[a] - lo[a], 1e-30f);
        int e = (int)ceilf(log2f(ext / 254.0f));
        e = max(-126, min(126, e));
        float s = __uint_as_float((uint32_t)(e + 127) << 23);
        while (s * 254.0f < ext && e < 126) { e++; s = __uint_as_float((uint32_t)(e + 127) << 23); }
        ebias[a] = e + 127;
        scale[a] = s;
    }

    // ---- greedy child -> slot assignment (octant order), Ylitie et al. section 3.2
    int slot_of[8];
    {
        const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
        float costm[8][8];
        for (int c = 0; c < nc; c++) {
            const float dx = 0.5f * (clo[c][0] + chi[c][0]) - cx;
            const float dy = 0.5f * (clo[c][1] + chi[c][1]) - cy;
            const float dz = 0.5f * (clo[c][2] + chi[c][2]) - cz;
            for (int s = 0; s < 8; s++)
                costm[c][s] = ((s & 1) ? dx : -dx) + ((s & 2) ? dy : -dy) + ((s & 4) ? dz : -dz);
        }
        bool c_done[8] = {false, false, false, false, false, false, false, false};
        bool s_done[8] = {false, false, false, false, false, false, false, false};
        for (int r = 0; r < nc; r++) {
            float best = -FLT_MAX;
            int bc = -1, bs = -1;
            for (int c = 0; c < nc; c++) {
                if (c_done[c]) continue;
                for (int s = 0; s < 8; s++) {
                    if (s_done[s]) continue;
                    if (costm[c][s] > best) { best = costm[c][s]; bc = c; bs = s; }
                }
            }
            c_done[bc] = true; s_done[bs] = true; slot_of[bc] = bs;
        }
    }
    int child_in_slot[8];
    for (int s = 0; s < 8; s++) child_in_slot[s] = -1;
    for (int c = 0; c < nc; c++) child_in_slot[slot_of[c]] = c;

    // ---- allocate children + triangles
    int n_inner = 0, n_tri = 0;
    for (int c = 0; c < nc; c++) { if (c_inner[c]) n_inner++; else n_tri += c_cnt[c]; }
    const uint32_t child_base = n_inner ? atomicAdd(&ctr->n_wide, (uint32_t)n_inner) : 0u;
    const uint32_t tri_base = n_tri ? atomicAdd(&ctr->n_tris, (uint32_t)n_tri) : 0u;
    const uint32_t next_base = n_inner ? atomicAdd(&ctr->n_next, (uint32_t)n_inner) : 0u;
    atomicMax(&ctr->max_depth, (uint32_t)it.depth);

    uint32_t imask = 0, tword = 0;   // tword: bits 3s..3s+cnt-1 set for a leaf with cnt triangles in slot s (traverse.cuh)
    uint8_t qlo[3][8], qhi[3][8];
    int inner_rank = 0, tri_off = 0;
    float sah = 0.f;
    for (int s = 0; s < 8; s++) {
        const int c = child_in_slot[s];
        for (int a = 0; a < 3; a++) { qlo[a][s] = 255; qhi[a][s] = 0; }   // empty slot: an inverted box, never entered
        if (c < 0) continue;
        for (int a = 0; a < 3; a++) {
            float ql = floorf((clo[c][a] - lo[a]) / scale[a]);
            ql = fminf(fmaxf(ql, 0.f), 255.f);
            while (ql > 0.f && !(__fadd_ru(lo[a], ql * scale[a]) <= clo[c][a])) ql -= 1.f;
            float qh = ceilf((chi[c][a] - lo[a]) / scale[a]);
            qh = fminf(fmaxf(qh, 0.f), 255.f);
            while (qh < 255.f && !(__fadd_rd(lo[a], qh * scale[a]) >= chi[c][a])) qh += 1.f;
            qlo[a][s] = (uint8_t)ql;
            qhi[a][s] = (uint8_t)qh;
        }
        const float4 l4 = make_float4(clo[c][0], clo[c][1], clo[c][2], 0.f), h4 = make_float4(chi[c][0], chi[c][1], chi[c][2], 0.f);
        if (c_inner[c]) {
            imask |= 1u << s;
            next[next_base + inner_rank] = EmitItem{c_node[c], (int)(child_base + inner_rank), it.depth + 1};
            inner_rank++;
            sah += half_area(l4, h4) * kCostNode;
        } else {
            const int cnt = c_cnt[c];
            tword |= ((1u << cnt) - 1u) << (3 * s);   // its triangles follow those of the lower slots: index = tri_base + rank in tword
            for (int t = 0; t < cnt; t++) {
                const uint32_t prim = sorted_prim[c_first[c] + t];
                const float4 a = tri_pos[3 * (size_t)prim], b = tri_pos[3 * (size_t)prim + 1], cc = tri_pos[3 * (size_t)prim + 2];
                const uint32_t flags = (__float_as_int(b.w) >= 0) ? TRI_FLAG_SINGLE_SIDED : 0u;  // b.w = light id
                float4* o = out_tris + 3 * (size_t)(tri_base + tri_off + t);
                o[0] = make_float4(a.x, a.y, a.z, __uint_as_float(prim));
                o[1] = make_float4(__fsub_rn(b.x, a.x), __fsub_rn(b.y, a.y), __fsub_rn(b.z, a.z), __uint_as_float(flags));
                o[2] = make_float4(__fsub_rn(cc.x, a.x), __fsub_rn(cc.y, a.y), __fsub_rn(cc.z, a.z), 0.f);
            }
            tri_off += cnt;
            sah += half_area(l4, h4) * kCostPrim * cnt;
        }
    }
    atomicAdd(&ctr->sah, sah);

    auto pack4 = [](const uint8_t* b) { return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24); };
    float4* o = out_nodes + 5 * (size_t)it.wide;
    o[0] = make_float4(lo[0], lo[1], lo[2],
                       __uint_as_float((uint32_t)ebias[0] | ((uint32_t)ebias[1] << 8) | ((uint32_t)ebias[2] << 16) | (imask << 24)));
    o[1] = make_float4(__uint_as_float(child_base), __uint_as_float(tri_base), __uint_as_float(tword), 0.f);
    o[2] = make_float4(__uint_as_float(pack4(qlo[0])), __uint_as_float(pack4(qlo[0] + 4)), __uint_as_float(pack4(qlo[1])), __uint_as_float(pack4(qlo[1] + 4)));
    o[3] = make_float4(__uint_as_float(pack4(qlo[2])), __uint_as_float(pack4(qlo[2] + 4)), __uint_as_float(pack4(qhi[0])), __uint_as_float(pack4(qhi[0] + 4)));
    o[4] = make_float4(__uint_as_float(pack4(qhi[1])), __uint_as_float(pack4(qhi[1] + 4)), __uint_as_float(pack4(qhi[2])), __uint_as_float(pack4(qhi[2] + 4)));
}

}  // namespace

// Build scratch: one device allocation per build, carved up by a bump allocator (a dozen cudaMalloc/cudaFree pairs cost
// more than all the build kernels together: 9 -> 90 ms at 1 M triangles when every buffer was its own allocation).
struct BuildArena {
    char*  base = nullptr;
    size_t cap = 0, off = 0;
    ~BuildArena() { if (base) cudaFree(base); }
    void reserve(size_t bytes) {
        SPC_CUDA(cudaMalloc((void**)&base, bytes));
        cap = bytes;
    }
    void* take(size_t bytes) {
        const size_t at = (off + 255) & ~(size_t)255;
        SPC_REQUIRE(at + bytes <= cap, SPC_ERR_CAPACITY, "BVH build scratch exhausted (%zu + %zu of %zu bytes)", at, bytes, cap);
        off = at + bytes;
        return base + at;
    }
};
static BuildArena* g_arena = nullptr;   // valid inside build_bvh only (builds are serialised per process by g_build_mutex)
template <typename T>
struct ScratchBuf {
    T*     p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        if (count == 0) count = 1;
        p = (T*)g_arena->take(count * sizeof(T));
        n = count;
    }
    size_t bytes() const { return n * sizeof(T); }
};

static void stage_mark(cudaStream_t st, const char* what) {
    static const bool on = getenv("SPC_BVH_VERBOSE") != nullptr;
    if (!on) return;
    cudaStreamSynchronize(st);
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    fprintf(stderr, "[spc] build stage %-28s t=%.3f ms\n", what, (ts.tv_sec % 1000) * 1e3 + ts.tv_nsec * 1e-6);
}

void build_bvh(Context& ctx, const float4* d_tri_pos, uint32_t n) {
    NvtxRange range("spc: BVH build");
    SPC_REQUIRE(n >= 1, SPC_ERR_INVALID, "scene has no triangles");
    cudaStream_t st = ctx.stream;
    static std::mutex g_build_mutex;
    std::lock_guard<std::mutex> lock(g_build_mutex);
    cudaEvent_t ev0, ev1;
    SPC_CUDA(cudaEventCreate(&ev0));
    SPC_CUDA(cudaEventCreate(&ev1));
    SPC_CUDA(cudaEventRecord(ev0, st));
    {
        static const float cost_prim = []() { const char* e = getenv("SPC_BVH_COST_PRIM"); return e ? (float)atof(e) : 0.3f; }();
        SPC_CUDA(cudaMemcpyToSymbolAsync(g_cost_prim, &cost_prim, sizeof(float), 0, cudaMemcpyHostToDevice, st));
    }
    BuildArena arena;   // inside the timed region: build_ms includes the scratch allocation
    g_arena = &arena;
    arena.reserve((size_t)n * 420 + (64u << 20));
    stage_mark(st, "start");

    const int B = 256;
    const unsigned gN = (n + B - 1) / B;
    ScratchBuf<float4> plo, phi;
    plo.alloc(n); phi.alloc(n);
    ScratchBuf<uint32_t> scene;
    scene.alloc(6);
    {
        uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
        SPC_CUDA(cudaMemcpyAsync(scene.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    }
    k_prim_bounds<<<gN, B, 0, st>>>(d_tri_pos, n, plo.p, phi.p, scene.p);
    ctx.launches++;
    uint32_t hs[6];
    SPC_CUDA(cudaMemcpyAsync(hs, scene.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
    SPC_CUDA(cudaStreamSynchronize(st));
    float slo[3], shi[3], maxabs = 0.f;
    for (int a = 0; a < 3; a++) {
        slo[a] = ord2f(hs[a]);
        shi[a] = ord2f(hs[3 + a]);
        maxabs = fmaxf(maxabs, fmaxf(fabsf(slo[a]), fabsf(shi[a])));
        ctx.geom.scene_lo[a] = slo[a];
        ctx.geom.scene_hi[a] = shi[a];
    }
    SPC_REQUIRE(maxabs < 1e30f && maxabs == maxabs, SPC_ERR_INVALID, "scene has non-finite vertices");
    // conservative padding of leaf boxes: covers the rounding of the slab test and of the
    // triangle test (a few ulp of the coordinate magnitude); 2^-18 of the largest coordinate.
    const float pad = fmaxf(maxabs, 1e-20f) * (1.0f / 262144.0f);

    ScratchBuf<uint64_t> keys, keys2;
    ScratchBuf<uint32_t> vals, vals2;
    keys.alloc(n); keys2.alloc(n); vals.alloc(n); vals2.alloc(n);
    float3 fslo = make_float3(slo[0], slo[1], slo[2]);
    float3 sinv = make_float3(1.f / fmaxf(shi[0] - slo[0], 1e-30f), 1.f / fmaxf(shi[1] - slo[1], 1e-30f),
                              1.f / fmaxf(shi[2] - slo[2], 1e-30f));
    k_morton<<<gN, B, 0, st>>>(plo.p, phi.p, n, fslo, sinv, keys.p, vals.p);
    ctx.launches++;
    {
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys2.p, vals.p, vals2.p, (int)n, 0, 63, st);
        ScratchBuf<uint8_t> tmp;
        tmp.alloc(tmp_bytes);
        SPC_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys2.p, vals.p, vals2.p, (int)n, 0, 63, st));
        SPC_CUDA(cudaStreamSynchronize(st));
    }
    stage_mark(st, "morton + sort");
    const uint64_t* skeys = keys2.p;
    const uint32_t* sprim = vals2.p;

    const uint32_t n_int = n > 1 ? n - 1 : 0;
    ScratchBuf<int> left, right, parent, first, last, flags;
    left.alloc(n_int); right.alloc(n_int); first.alloc(n_int); last.alloc(n_int); flags.alloc(n_int);
    parent.alloc(2 * (size_t)n);
    ScratchBuf<float4> nb_lo, nb_hi;
    nb_lo.alloc(2 * (size_t)n); nb_hi.alloc(2 * (size_t)n);
    ScratchBuf<float> cost;
    ScratchBuf<uint8_t> dec;
    cost.alloc((size_t)n_int * 7);
    dec.alloc((size_t)n_int * 8);
    SPC_CUDA(cudaMemsetAsync(flags.p, 0, flags.bytes(), st));
    // binary topology: PLOC by default, SPC_BVH_BUILDER=lbvh selects the Morton radix tree (kept for A/B measurements)
    static const bool use_ploc = []() { const char* e = getenv("SPC_BVH_BUILDER"); return !(e && strcmp(e, "lbvh") == 0); }();
    ScratchBuf<int> parent2;
    ScratchBuf<uint32_t> sprim2;
    const int* parent_p = parent.p;
    int ploc_rounds = 0;
    if (n_int && use_ploc) {
        ScratchBuf<int> cidA, cidB, nn, cnt, leaf_pos;
        ScratchBuf<unsigned long long> fl, sc;
        cidA.alloc(n); cidB.alloc(n); nn.alloc(n); cnt.alloc(2 * (size_t)n); leaf_pos.alloc(n);
        fl.alloc(n); sc.alloc(n);
        size_t scan_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, fl.p, sc.p, (int)n, st);
        ScratchBuf<uint8_t> scan_tmp;
        scan_tmp.alloc(scan_bytes);
        stage_mark(st, "ploc allocs");
        k_ploc_init<<<gN, B, 0, st>>>((int)n, sprim, plo.p, phi.p, pad, nb_lo.p, nb_hi.p, cidA.p, cnt.p);
        ctx.launches++;
        stage_mark(st, "ploc init");
        int m = (int)n, merged = 0;
        int* ca = cidA.p;
        int* cb = cidB.p;
        const size_t nn_smem = (size_t)(B + 2 * kPlocRadius) * 2 * sizeof(float4);
        while (m > 1) {
            const unsigned g = (m + B - 1) / B;
            k_ploc_nn<<<g, B, nn_smem, st>>>(m, ca, nb_lo.p, nb_hi.p, nn.p);
            k_ploc_flags<<<g, B, 0, st>>>(m, nn.p, fl.p);
            SPC_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp.p, scan_bytes, fl.p, sc.p, m, st));
            k_ploc_apply<<<g, B, 0, st>>>(m, (int)n, merged, ca, nn.p, fl.p, sc.p, cb, left.p, right.p, parent.p, cnt.p, nb_lo.p, nb_hi.p);
            ctx.launches += 4;
            unsigned long long tail[2];
            SPC_CUDA(cudaMemcpyAsync(&tail[0], sc.p + (m - 1), sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            SPC_CUDA(cudaMemcpyAsync(&tail[1], fl.p + (m - 1), sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            SPC_CUDA(cudaStreamSynchronize(st));
            const unsigned long long tot = tail[0] + tail[1];
            const int kept = (int)(tot & 0xffffffffull), pairs = (int)(tot >> 32);
            SPC_REQUIRE(pairs > 0 && kept == m - pairs, SPC_ERR_CUDA, "PLOC round made no progress (%d clusters, %d pairs, %d kept)", m, pairs, kept);
            if (getenv("SPC_BVH_VERBOSE") && (ploc_rounds % 8 == 0 || m < 30)) {
                timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
                fprintf(stderr, "[spc] PLOC round %d: %d clusters, %d pairs  t=%.3f ms\n", ploc_rounds, m, pairs, ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6);
            }
            merged += pairs;
            m = kept;
            int* t = ca; ca = cb; cb = t;
            ploc_rounds++;
        }
        SPC_REQUIRE(merged == (int)n_int, SPC_ERR_CUDA, "PLOC built %d of %u internal nodes", merged, n_int);
        if (getenv("SPC_BVH_VERBOSE")) {
            SPC_CUDA(cudaEventRecord(ev1, st));
            SPC_CUDA(cudaStreamSynchronize(st));
            float t = 0.f;
            cudaEventElapsedTime(&t, ev0, ev1);
            fprintf(stderr, "[spc] PLOC: %d rounds for %u triangles, %.2f ms since build start\n", ploc_rounds, n, t);
        }
        const int root_parent = -1;
        SPC_CUDA(cudaMemcpyAsync(parent.p, &root_parent, sizeof(int), cudaMemcpyHostToDevice, st));
        parent2.alloc(2 * (size_t)n);
        sprim2.alloc(n);
        const unsigned g2 = (2 * n - 1 + B - 1) / B;
        k_ploc_order<<<g2, B, 0, st>>>((int)n, left.p, parent.p, cnt.p, sprim, sprim2.p, first.p, last.p, leaf_pos.p);
        k_ploc_relink<<<g2, B, 0, st>>>((int)n, leaf_pos.p, left.p, right.p, parent.p, parent2.p);
        ctx.launches += 2;
        SPC_CUDA(cudaStreamSynchronize(st));   // the round buffers go out of scope here
        stage_mark(st, "ploc order + relink");
        sprim = sprim2.p;
        parent_p = parent2.p;
    } else if (n_int) {
        k_radix_tree<<<(n_int + B - 1) / B, B, 0, st>>>(skeys, (int)n, left.p, right.p, parent.p, first.p, last.p);
        ctx.launches++;
    }
    k_fit_dp<<<gN, B, 0, st>>>((int)n, sprim, plo.p, phi.p, pad, left.p, right.p, parent_p, first.p, last.p,
                              nb_lo.p, nb_hi.p, cost.p, dec.p, flags.p);
    ctx.launches++;
    SPC_CUDA(cudaGetLastError());

    stage_mark(st, "fit + collapse DP");
    // emission, level by level
    const size_t max_wide = (size_t)(n_int ? n_int : 1);
    ScratchBuf<float4> tmp_nodes;
    tmp_nodes.alloc(max_wide * 5);
    ctx.bvh.tris.alloc((size_t)n * 3);
    ScratchBuf<EmitItem> qa, qb;
    qa.alloc(max_wide); qb.alloc(max_wide);
    ScratchBuf<EmitCounters> ctr;
    ctr.alloc(1);
    EmitCounters hc = {1u, 0u, 0u, 0u, 0.f};
    SPC_CUDA(cudaMemcpyAsync(ctr.p, &hc, sizeof(hc), cudaMemcpyHostToDevice, st));
    EmitItem root = {n_int ? 0 : 0, 0, 1};
    SPC_CUDA(cudaMemcpyAsync(qa.p, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    int n_items = 1;
    EmitItem* cur = qa.p;
    EmitItem* nxt = qb.p;
    while (n_items > 0) {
        k_emit<<<(n_items + 63) / 64, 64, 0, st>>>((int)n, cur, n_items, nxt, ctr.p, left.p, right.p, first.p, last.p,
                                                dec.p, nb_lo.p, nb_hi.p, sprim, d_tri_pos, tmp_nodes.p, ctx.bvh.tris.p);
        ctx.launches++;
        SPC_CUDA(cudaGetLastError());
        SPC_CUDA(cudaMemcpyAsync(&hc, ctr.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
        SPC_CUDA(cudaStreamSynchronize(st));
        n_items = (int)hc.n_next;
        hc.n_next = 0;
        SPC_CUDA(cudaMemcpyAsync(&ctr.p->n_next, &hc.n_next, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        EmitItem* t = cur; cur = nxt; nxt = t;
    }
    stage_mark(st, "emit");
    SPC_REQUIRE(hc.n_tris == n, SPC_ERR_CUDA, "BVH emission lost triangles: %u of %u", hc.n_tris, n);
    SPC_REQUIRE((int)hc.max_depth <= kMaxBvhDepth, SPC_ERR_CAPACITY, "BVH depth %u exceeds traversal stack %d",
                hc.max_depth, kMaxBvhDepth);
    ctx.bvh.n_nodes = hc.n_wide;
    ctx.bvh.n_tris = n;
    ctx.bvh.nodes.alloc((size_t)hc.n_wide * 5);
    SPC_CUDA(cudaMemcpyAsync(ctx.bvh.nodes.p, tmp_nodes.p, (size_t)hc.n_wide * 80, cudaMemcpyDeviceToDevice, st));

    float4 rl, rh;
    const int root_id = n_int ? 0 : 0;
    SPC_CUDA(cudaMemcpyAsync(&rl, nb_lo.p + root_id, sizeof(float4), cudaMemcpyDeviceToHost, st));
    SPC_CUDA(cudaMemcpyAsync(&rh, nb_hi.p + root_id, sizeof(float4), cudaMemcpyDeviceToHost, st));
    SPC_CUDA(cudaEventRecord(ev1, st));
    SPC_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    SPC_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    const float ra = fmaxf((rh.x - rl.x) * (rh.y - rl.y) + (rh.y - rl.y) * (rh.z - rl.z) + (rh.z - rl.z) * (rh.x - rl.x), 1e-30f);
    spc_bvh_stats& s = ctx.bvh_stats;
    s.n_triangles = n;
    s.n_nodes = hc.n_wide;
    s.n_bvh2_nodes = 2 * n - 1;
    s.max_depth = hc.max_depth;
    s.sah_cost = kCostNode + hc.sah / ra;
    s.build_ms = ms;
    s.bytes_nodes = (uint64_t)hc.n_wide * 80;
    s.bytes_triangles = (uint64_t)n * 48;
}

}  // namespace spc
