// traverse.cuh -- device-side traversal of the compressed 8-wide BVH (replaces optixTrace,
// reference call sites cuProg.h:395,420,445 (closest hit) and cuProg.h:470 (occlusion)).
//
// Intersection contract (must stay bit-identical to oracle/spc_oracle.cpp: orc_tri_test):
//   dot(a,b)   = fma(a.z,b.z, fma(a.y,b.y, a.x*b.x))
//   cross(a,b) = ( fma(a.y,b.z, -(a.z*b.y)), fma(a.z,b.x, -(a.x*b.z)), fma(a.x,b.y, -(a.y*b.x)) )
//   e1 = v1-v0, e2 = v2-v0;  pvec = cross(d,e2);  det = dot(e1,pvec)
//   miss when det == 0 (single-sided triangles under CULL_BACK_FACING: miss unless det > 0)
//   inv = 1/det; tvec = o-v0; u = dot(tvec,pvec)*inv; miss unless 0 <= u <= 1
//   qvec = cross(tvec,e1); v = dot(d,qvec)*inv; miss unless v >= 0 and u+v <= 1
//   t = dot(e2,qvec)*inv;  hit when tmin < t < tmax; nearest t wins, equal t -> lowest prim id.
// Every operation is a single IEEE-754 binary32 op (explicit intrinsics: no contraction, no ftz).
// The result is therefore independent of BVH shape and traversal order.
#pragma once
#include "common.cuh"

namespace spc {

constexpr int kSmStack  = 8;    // per-thread stack entries kept in shared memory
constexpr int kLocStack = 24;   // overflow entries in local memory
constexpr int kMaxBvhDepth = kSmStack + kLocStack;

__device__ __forceinline__ float c_dot(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
}
__device__ __forceinline__ void c_cross(float ax, float ay, float az, float bx, float by, float bz,
                                        float& cx, float& cy, float& cz) {
    cx = __fmaf_rn(ay, bz, -__fmul_rn(az, by));
    cy = __fmaf_rn(az, bx, -__fmul_rn(ax, bz));
    cz = __fmaf_rn(ax, by, -__fmul_rn(ay, bx));
}

// byte j of w, biased: the float 2^23 + 256*byte, built by splicing the byte into mantissa bits 8..15 of 2^23 (one PRMT, no
// I2F and no subtraction).  The bias is folded into the per-node plane origins (trav_step), so a child plane costs PRMT + FFMA.
// The bias word travels in a register (trav_bias()): PRMT takes one immediate, and with the selector as the immediate
// ptxas no longer re-materialises the four selectors in front of every PRMT (48 moves per node in the r1d SASS).
template <int J>
__device__ __forceinline__ float byte_to_biased_float(uint32_t w, uint32_t bias) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(bias), "n"(0x7604 + 16 * J));
    return __uint_as_float(r);
}
// 0x4B000000 (the float 2^23) as a value ptxas cannot fold into an immediate: gridDim.z is 1 for every launch of this library.
__device__ __forceinline__ uint32_t trav_bias() { return 0x4A000000u + (gridDim.z << 24); }
// per-thread traversal stack in shared memory, addressed by 32-bit shared-window offsets
__device__ __forceinline__ void sm_push(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint2 sm_pop(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t sign_extend_s8x4(uint32_t x) {
    uint32_t r;
    asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int j) { return (w >> (8 * j)) & 0xffu; }

// Two small look-up tables in shared memory turn the 8-bit "which child boxes were hit" mask of a node into the two things the
// traversal needs, instead of shifting a per-child meta byte into place for each of the 8 children (5 predicated ALU operations per
// child in the r1d/r1e SASS; the kernel is bound by the half-rate ALU pipe):
//   exp3[m]     bit j of m -> bits 3j..3j+2: ANDed with the node's triangle word T (bits 3j..3j+cnt-1 set for a leaf in slot j)
//               it gives the triangles of the hit leaves; a triangle's index is tri_base + its rank among the set bits of T
//   perm[o][x]  bit j of x -> bit j^o: the hit inner children in traversal priority order for ray octant o (Ylitie et al. 2017)
struct TravLut {
    uint32_t exp3[256];
    uint8_t  perm[8 * 256];
};
// every kernel that traverses calls this once, before any thread leaves (it ends in __syncthreads)
__device__ __forceinline__ void trav_lut_init(TravLut& L) {
    for (int m = threadIdx.x; m < 256; m += blockDim.x) {
        uint32_t e = 0;
        for (int j = 0; j < 8; j++)
            if (m & (1 << j)) e |= 7u << (3 * j);
        L.exp3[m] = e;
    }
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) {
        const int o = i >> 8, x = i & 255;
        uint32_t r = 0;
        for (int j = 0; j < 8; j++)
            if (x & (1 << j)) r |= 1u << (j ^ o);
        L.perm[i] = (uint8_t)r;
    }
    __syncthreads();
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

struct TravRay {
    float ox, oy, oz, dx, dy, dz, tmin, tmax;
};
struct TravHit {
    float t, u, v;
    int   prim;
};

// Traversal state of one ray.  The loop body is exposed as trav_step() so that the persistent wavefront kernels
// (trace.cu) can interleave ray fetches with traversal steps; traverse_bvh8() runs it to completion for the
// megakernel-style passes (light trace, training tracer, pt).
struct Trav {
    float ox, oy, oz, dx, dy, dz, tmin, tmax;
    float idx, idy, idz;
    uint32_t oct_inv;
    float tcur, best_u, best_v;
    int   best_prim;
    uint2 ngroup;
    int   sp;
    uint32_t bias;    // trav_bias()
    uint32_t sbase;   // shared-window address of this lane's stack slot 0
};

__device__ __forceinline__ void trav_init(Trav& s, const TravRay& r) {
    const float eps = 1.0e-24f;
    s.ox = r.ox; s.oy = r.oy; s.oz = r.oz; s.dx = r.dx; s.dy = r.dy; s.dz = r.dz; s.tmin = r.tmin; s.tmax = r.tmax;
    s.idx = 1.0f / (fabsf(r.dx) > eps ? r.dx : copysignf(eps, r.dx));
    s.idy = 1.0f / (fabsf(r.dy) > eps ? r.dy : copysignf(eps, r.dy));
    s.idz = 1.0f / (fabsf(r.dz) > eps ? r.dz : copysignf(eps, r.dz));
    const uint32_t oct = (r.dx < 0.f ? 1u : 0u) | (r.dy < 0.f ? 2u : 0u) | (r.dz < 0.f ? 4u : 0u);
    s.oct_inv = 7u ^ oct;
    s.tcur = r.tmax;
    s.best_prim = -1;
    s.best_u = s.best_v = 0.f;
    s.ngroup = make_uint2(0u, 0x80000000u);
    s.sp = 0;
    s.bias = trav_bias();
}

// one child quad (4 of the 8 slots) of a node
#define SPC_CHILD_TEST2(J, SLOT)                                                                  \
    {                                                                                             \
        float lx = __fmaf_rn(byte_to_biased_float<J>(slox, bias), adjx, olx);                     \
        float ly = __fmaf_rn(byte_to_biased_float<J>(sloy, bias), adjy, oly);                     \
        float lz = __fmaf_rn(byte_to_biased_float<J>(sloz, bias), adjz, olz);                     \
        float hx = __fmaf_rn(byte_to_biased_float<J>(shix, bias), adjx, ohx);                     \
        float hy = __fmaf_rn(byte_to_biased_float<J>(shiy, bias), adjy, ohy);                     \
        float hz = __fmaf_rn(byte_to_biased_float<J>(shiz, bias), adjz, ohz);                     \
        float cmin = fmaxf(fmaxf(lx, ly), fmaxf(lz, s.tmin));                                     \
        float cmax = fminf(fminf(hx, hy), fminf(hz, s.tcur));                                     \
        if (cmin <= cmax) hit8 |= 1u << (SLOT);                                                   \
    }

// One iteration: take the nearest pending child of the current node group (or fall back to its triangles), test
// its 8 children, intersect the triangles of the hit leaves, pop when the group is exhausted.
// Returns true when the ray is finished (stack empty, or first hit for ANYHIT -- then s.best_prim >= 0).
// POSTPONE (wavefront kernels): when fewer than 1/5 of the lanes that entered the triangle loop are still in it, the stragglers
// push their remaining triangles back on the stack and rejoin the warp for the next node step (Ylitie et al. 2017, "triangle
// postponing").  The result does not depend on the order triangles are tested in (intersection contract above).
// current triangle group: node index, hit triangles (bits of the node's triangle word), first triangle, triangle word
struct TriGroup {
    uint32_t node, mask, base, T;
};

__device__ __forceinline__ void trav_push(Trav& s, uint32_t sstride_b, uint2* lstack, uint2 v) {
    if (s.sp < kSmStack) sm_push(s.sbase + s.sp * sstride_b, v);
    else lstack[s.sp - kSmStack] = v;
    s.sp++;
}
__device__ __forceinline__ uint2 trav_pop(Trav& s, uint32_t sstride_b, uint2* lstack) {
    s.sp--;
    return (s.sp < kSmStack) ? sm_pop(s.sbase + s.sp * sstride_b) : lstack[s.sp - kSmStack];
}

// Node step (requires s.ngroup.y > 0x00ffffff): take the nearest pending child of the current node group, test its 8 children;
// s.ngroup becomes that node's hit inner children, g the triangles of its hit leaves.
template <bool COUNT>
__device__ __forceinline__ void trav_node(const float4* __restrict__ nodes, Trav& s, uint32_t sstride_b, uint2* lstack, uint32_t lut, TriGroup& g,
                                          unsigned& cnt_nodes) {
    const uint32_t oct_inv = s.oct_inv;
    const uint32_t oct = 7u ^ oct_inv;
    const uint32_t bias = s.bias;
    const uint32_t hits  = s.ngroup.y;
    const uint32_t imask = s.ngroup.y & 0xffu;
    const uint32_t bit   = 31u - __clz(hits);
    s.ngroup.y &= ~(1u << bit);
    if (s.ngroup.y > 0x00ffffffu) {
        trav_push(s, sstride_b, lstack, s.ngroup);
    }
    const uint32_t slot = (bit - 24u) ^ oct_inv;
    const uint32_t rel  = __popc(imask & ~(0xffffffffu << slot));
    g.node = s.ngroup.x + rel;
    const float4*  np   = nodes + (size_t)g.node * 5;
    const float4 n0 = __ldg(np + 0);
    const float4 n1 = __ldg(np + 1);
    const float4 n2 = __ldg(np + 2);
    const float4 n3 = __ldg(np + 3);
    const float4 n4 = __ldg(np + 4);
    if (COUNT) cnt_nodes++;

    const uint32_t e_im = __float_as_uint(n0.w);
    // Child plane q (a byte) lies at t = q*step*id + (origin - o)*id.  The byte arrives as B = 2^23 + 256 q, so with
    // adj = step*id/256 the plane is t = B*adj + (org - 2^23*adj): one FFMA per plane.  Folding the bias costs at most
    // |adj|/2 of rounding in the constant (1/512 of a quantisation step, when the node is near the ray origin); entry planes
    // are moved one |adj| (1/256 step) earlier and exit planes one |adj| later, which more than covers it -- the box test
    // stays conservative and the hit set (decided by the triangle contract alone) is unchanged.
    const float adjx = __uint_as_float((e_im & 0xffu) << 23) * s.idx * 0.00390625f;
    const float adjy = __uint_as_float(((e_im >> 8) & 0xffu) << 23) * s.idy * 0.00390625f;
    const float adjz = __uint_as_float(((e_im >> 16) & 0xffu) << 23) * s.idz * 0.00390625f;
    const float orgx = __fmaf_rn(-8388608.0f, adjx, (n0.x - s.ox) * s.idx);
    const float orgy = __fmaf_rn(-8388608.0f, adjy, (n0.y - s.oy) * s.idy);
    const float orgz = __fmaf_rn(-8388608.0f, adjz, (n0.z - s.oz) * s.idz);
    const float olx = orgx - fabsf(adjx), ohx = orgx + fabsf(adjx);
    const float oly = orgy - fabsf(adjy), ohy = orgy + fabsf(adjy);
    const float olz = orgz - fabsf(adjz), ohz = orgz + fabsf(adjz);

    uint32_t hit8 = 0;
    {
        const uint32_t qlox = __float_as_uint(n2.x), qloy = __float_as_uint(n2.z), qloz = __float_as_uint(n3.x);
        const uint32_t qhix = __float_as_uint(n3.z), qhiy = __float_as_uint(n4.x), qhiz = __float_as_uint(n4.z);
        const uint32_t slox = (oct & 1u) ? qhix : qlox, shix = (oct & 1u) ? qlox : qhix;
        const uint32_t sloy = (oct & 2u) ? qhiy : qloy, shiy = (oct & 2u) ? qloy : qhiy;
        const uint32_t sloz = (oct & 4u) ? qhiz : qloz, shiz = (oct & 4u) ? qloz : qhiz;
        SPC_CHILD_TEST2(0, 0) SPC_CHILD_TEST2(1, 1) SPC_CHILD_TEST2(2, 2) SPC_CHILD_TEST2(3, 3)
    }
    {
        const uint32_t qlox = __float_as_uint(n2.y), qloy = __float_as_uint(n2.w), qloz = __float_as_uint(n3.y);
        const uint32_t qhix = __float_as_uint(n3.w), qhiy = __float_as_uint(n4.y), qhiz = __float_as_uint(n4.w);
        const uint32_t slox = (oct & 1u) ? qhix : qlox, shix = (oct & 1u) ? qlox : qhix;
        const uint32_t sloy = (oct & 2u) ? qhiy : qloy, shiy = (oct & 2u) ? qloy : qhiy;
        const uint32_t sloz = (oct & 4u) ? qhiz : qloz, shiz = (oct & 4u) ? qloz : qhiz;
        SPC_CHILD_TEST2(0, 4) SPC_CHILD_TEST2(1, 5) SPC_CHILD_TEST2(2, 6) SPC_CHILD_TEST2(3, 7)
    }
    const uint32_t imask_new = e_im >> 24;
    g.T = __float_as_uint(n1.z);
    g.mask = g.T & lds_u32(lut + 4u * hit8);
    g.base = __float_as_uint(n1.y);
    const uint32_t ph = lds_u8(lut + 1024u + (oct_inv << 8) + (hit8 & imask_new));
    s.ngroup.x = __float_as_uint(n1.x);
    s.ngroup.y = (ph << 24) | imask_new;
}

// a parked triangle group {node index, triangle mask} sits in s.ngroup: its base and triangle word are re-read from the node
__device__ __forceinline__ void trav_parked(const float4* __restrict__ nodes, Trav& s, TriGroup& g) {
    g.node = s.ngroup.x;
    g.mask = s.ngroup.y;
    const float4 n1 = __ldg(nodes + (size_t)g.node * 5 + 1);
    g.base = __float_as_uint(n1.y);
    g.T = __float_as_uint(n1.z);
    s.ngroup = make_uint2(0u, 0u);
}

// Triangles of g under the intersection contract.  Returns true when ANYHIT found an occluder (s.best_prim >= 0).
// POSTPONE (wavefront kernels): when fewer than 1/postpone_div of the lanes that entered the loop are still in it, the stragglers
// park their remaining triangles on the stack and rejoin the warp (Ylitie et al. 2017, "triangle postponing").  The result does not
// depend on the order triangles are tested in.
template <bool ANYHIT, bool COUNT, bool POSTPONE>
__device__ __forceinline__ bool trav_tris(const float4* __restrict__ tris, Trav& s, TriGroup& g, bool cull_back, uint32_t sstride_b, uint2* lstack,
                                          unsigned& cnt_tris, int postpone_div) {
    const int tri_lanes = POSTPONE ? __popc(__activemask()) : 0;
    while (g.mask != 0u) {
        if (POSTPONE && __popc(__activemask()) * postpone_div < tri_lanes) {
            trav_push(s, sstride_b, lstack, make_uint2(g.node, g.mask));
            break;
        }
        const uint32_t tb = 31u - __clz(g.mask);
        g.mask &= ~(1u << tb);
        const uint32_t ti = __popc(g.T & ~(0xffffffffu << tb));   // rank of this triangle among the node's triangles
        const float4* tp = tris + (size_t)(g.base + ti) * 3;
        const float4 a = __ldg(tp + 0);
        const float4 b = __ldg(tp + 1);
        const float4 c = __ldg(tp + 2);
        if (COUNT) cnt_tris++;
        float px, py, pz;
        c_cross(s.dx, s.dy, s.dz, c.x, c.y, c.z, px, py, pz);
        const float det = c_dot(b.x, b.y, b.z, px, py, pz);
        const bool  single = cull_back && (__float_as_uint(b.w) & TRI_FLAG_SINGLE_SIDED);
        if (single ? !(det > 0.0f) : !(det != 0.0f)) continue;
        const float inv = __fdiv_rn(1.0f, det);
        const float tx = __fsub_rn(s.ox, a.x), ty = __fsub_rn(s.oy, a.y), tz = __fsub_rn(s.oz, a.z);
        const float u = __fmul_rn(c_dot(tx, ty, tz, px, py, pz), inv);
        if (!(u >= 0.0f && u <= 1.0f)) continue;
        float qx, qy, qz;
        c_cross(tx, ty, tz, b.x, b.y, b.z, qx, qy, qz);
        const float v = __fmul_rn(c_dot(s.dx, s.dy, s.dz, qx, qy, qz), inv);
        if (!(v >= 0.0f && __fadd_rn(u, v) <= 1.0f)) continue;
        const float t = __fmul_rn(c_dot(c.x, c.y, c.z, qx, qy, qz), inv);
        if (!(t > s.tmin)) continue;
        const int prim = (int)__float_as_uint(a.w);
        if (ANYHIT) {
            if (t < s.tmax) {
                s.best_prim = prim;
                return true;
            }
        } else {
            if (t < s.tcur || (t == s.tcur && prim < s.best_prim)) {
                s.tcur = t;
                s.best_prim = prim;
                s.best_u = u;
                s.best_v = v;
            }
        }
    }
    return false;
}

// One iteration: node step (or a parked triangle group), the triangles of the hit leaves, pop when the group is exhausted.
// Returns true when the ray is finished (stack empty, or first hit for ANYHIT -- then s.best_prim >= 0).
template <bool ANYHIT, bool COUNT, bool POSTPONE = false>
__device__ __forceinline__ bool trav_step(const float4* __restrict__ nodes, const float4* __restrict__ tris, Trav& s, bool cull_back,
                                          int sstride, uint2* lstack, unsigned& cnt_nodes, unsigned& cnt_tris, uint32_t lut, int postpone_div = 5) {
    const uint32_t sstride_b = (uint32_t)sstride * 8u;
    TriGroup g;
    if (s.ngroup.y > 0x00ffffffu) trav_node<COUNT>(nodes, s, sstride_b, lstack, lut, g, cnt_nodes);
    else trav_parked(nodes, s, g);
    if (trav_tris<ANYHIT, COUNT, POSTPONE>(tris, s, g, cull_back, sstride_b, lstack, cnt_tris, postpone_div)) return true;
    if (s.ngroup.y <= 0x00ffffffu) {
        if (s.sp == 0) return true;
        s.ngroup = trav_pop(s, sstride_b, lstack);
    }
    return false;
}

// SPC_NODE_STEPS node steps per triangle phase (persistent wavefront kernels).  The triangle loop is where the warp is emptiest (9 of 32 lanes:
// only the lanes whose node had hit leaves take part, profiles/r1e_summary.md); collecting the leaves of two consecutive node steps
// before intersecting roughly doubles its occupancy and halves the number of passes.  A lane that gets triangles from both steps
// parks the older group on its stack.  Order of node visits and triangle tests changes, results do not (intersection contract).
#ifndef SPC_NODE_STEPS
#define SPC_NODE_STEPS 3   // node steps per triangle phase: 1 -> 2 -> 3 -> 4 gave 3776 / 4057 / 4150 / 4153 Mrays/s on bench.py (profiles/r1e_summary.md)
#endif
#ifndef SPC_NODE_STEPS_ANYHIT
#define SPC_NODE_STEPS_ANYHIT 3   // occlusion rays: steps before the triangles delay the early exit, yet 3 wins on the house scene too
                                  // (shadow-ray kernel time of two frames: 5.47 / 4.95 / 4.85 ms for 1 / 2 / 3 steps, tests/quick_anyhit_steps.sh)
#endif
template <bool ANYHIT, bool COUNT>
__device__ __forceinline__ bool trav_step2(const float4* __restrict__ nodes, const float4* __restrict__ tris, Trav& s, bool cull_back,
                                           int sstride, uint2* lstack, uint32_t lut, int postpone_div, unsigned& cn, unsigned& ct) {
    const uint32_t sstride_b = (uint32_t)sstride * 8u;
    // here s.ngroup.y == 0 means "no group" (a node group whose inner hits are used up keeps its imask in y: normalised below)
    if (s.ngroup.y == 0u) s.ngroup = trav_pop(s, sstride_b, lstack);   // sp > 0: otherwise the previous call returned true
    TriGroup p;
    if (s.ngroup.y > 0x00ffffffu) {
        trav_node<COUNT>(nodes, s, sstride_b, lstack, lut, p, cn);
        if (s.ngroup.y <= 0x00ffffffu) s.ngroup.y = 0u;
    } else {
        trav_parked(nodes, s, p);
    }
    constexpr int kSteps = ANYHIT ? SPC_NODE_STEPS_ANYHIT : SPC_NODE_STEPS;
#pragma unroll
    for (int k = 1; k < kSteps; k++) {
        if (s.ngroup.y == 0u && s.sp > 0) s.ngroup = trav_pop(s, sstride_b, lstack);
        if (s.ngroup.y > 0x00ffffffu) {   // (a parked triangle group popped here waits for the next call)
            TriGroup q;
            trav_node<COUNT>(nodes, s, sstride_b, lstack, lut, q, cn);
            if (s.ngroup.y <= 0x00ffffffu) s.ngroup.y = 0u;
            if (q.mask != 0u) {
                if (p.mask != 0u) trav_push(s, sstride_b, lstack, make_uint2(p.node, p.mask));
                p = q;
            }
        }
    }
    if (trav_tris<ANYHIT, COUNT, true>(tris, s, p, cull_back, sstride_b, lstack, ct, postpone_div)) return true;
    return s.ngroup.y == 0u && s.sp == 0;
}

template <bool ANYHIT, bool COUNT>
__device__ __forceinline__ bool traverse_bvh8(const float4* __restrict__ nodes,
                                              const float4* __restrict__ tris, const TravRay& r,
                                              bool cull_back, uint2* sstack, int sstride,
                                              TravHit& hit, unsigned& cnt_nodes, unsigned& cnt_tris, const TravLut& lut) {
    Trav s;
    trav_init(s, r);
    s.sbase = (uint32_t)__cvta_generic_to_shared(sstack);
    const uint32_t lut_addr = (uint32_t)__cvta_generic_to_shared(&lut);
    uint2 lstack[kLocStack];
    while (!trav_step<ANYHIT, COUNT>(nodes, tris, s, cull_back, sstride, lstack, cnt_nodes, cnt_tris, lut_addr)) {}
    if (ANYHIT) return s.best_prim >= 0;
    hit.t = s.best_prim >= 0 ? s.tcur : 0.0f;
    hit.u = s.best_u;
    hit.v = s.best_v;
    hit.prim = s.best_prim;
    return s.best_prim >= 0;
}

}  // namespace spc
