// trace.cu -- wavefront ray-batch kernels: closest hit and occlusion over the compressed 8-wide BVH.
// Stand in for the optixTrace calls of the reference (cuProg.h:384-487); one lane per ray, rays
// and hits are 128-bit coalesced loads/stores (32 B in, 16 B out), traversal stacks live in shared
// memory (kSmStack entries per lane) with a local-memory tail.
#include <algorithm>
#include <cstdlib>
#include "geom.cuh"
#include "traverse.cuh"

namespace spc {

constexpr int kTraceBlock = 128;
#ifndef SPC_PERSIST_MIN_BLOCKS
#define SPC_PERSIST_MIN_BLOCKS 8   // resident blocks per SM the persistent kernels are compiled for (register cap 65536 / (128 * n))
#endif

template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
k_trace_closest(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                const float4* __restrict__ rays, int64_t n, int cull_back, float4* __restrict__ hits,
                unsigned long long* __restrict__ counters) {
    __shared__ uint2 s_stack[kSmStack * kTraceBlock];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);
    const int64_t i = (int64_t)blockIdx.x * kTraceBlock + threadIdx.x;
    unsigned cn = 0, ct = 0;
    if (i < n) {
        const float4 ro = __ldg(rays + 2 * i);
        const float4 rd = __ldg(rays + 2 * i + 1);
        TravRay r{ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w, rd.w};
        TravHit h;
        traverse_bvh8<false, COUNT>(nodes, tris, r, cull_back != 0, s_stack + threadIdx.x, kTraceBlock, h, cn, ct, s_lut);
        hits[i] = make_float4(h.t, h.u, h.v, __int_as_float(h.prim));
    }
    if (COUNT) {
        // warp-aggregate, then one atomic pair per warp
        for (int o = 16; o > 0; o >>= 1) {
            cn += __shfl_xor_sync(0xffffffffu, cn, o);
            ct += __shfl_xor_sync(0xffffffffu, ct, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(counters + 0, (unsigned long long)cn);
            atomicAdd(counters + 1, (unsigned long long)ct);
        }
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
k_trace_occlusion(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                  const float4* __restrict__ rays, int64_t n, uint8_t* __restrict__ visible,
                  unsigned long long* __restrict__ counters) {
    __shared__ uint2 s_stack[kSmStack * kTraceBlock];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);
    const int64_t i = (int64_t)blockIdx.x * kTraceBlock + threadIdx.x;
    unsigned cn = 0, ct = 0;
    if (i < n) {
        const float4 ro = __ldg(rays + 2 * i);
        const float4 rd = __ldg(rays + 2 * i + 1);
        TravRay r{ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w, rd.w};
        TravHit h;
        const bool blocked = traverse_bvh8<true, COUNT>(nodes, tris, r, false, s_stack + threadIdx.x, kTraceBlock, h, cn, ct, s_lut);
        visible[i] = blocked ? 0 : 1;
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            cn += __shfl_xor_sync(0xffffffffu, cn, o);
            ct += __shfl_xor_sync(0xffffffffu, ct, o);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(counters + 0, (unsigned long long)cn);
            atomicAdd(counters + 1, (unsigned long long)ct);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Persistent warps with dynamic ray fetch.  A plain one-ray-per-lane launch leaves a warp running until its longest
// ray ends (measured: 6.7 of 32 lanes active per issued instruction on incoherent rays, profiles/r1a_summary.md).
// Here each warp keeps pulling rays from a global counter: whenever at least `fetch_threshold` lanes are idle the
// idle lanes claim new rays with one warp-aggregated atomicAdd, so lanes stay busy until the batch is drained.
// The grid is sized to the resident capacity of the GPU (SM count x blocks per SM), not to the ray count.
// ---------------------------------------------------------------------------------------------
// RAYGEN: ray i is the camera ray of pixel i, generated here (seed tea<4>(i, sample_index), jitter, camera_dir_exact -- the very
// arithmetic k_eye_shade repeats for the first bounce) instead of being read from a ray buffer another kernel would have to write.
template <bool ANYHIT, bool COUNT = false, bool RAYGEN = false>
__global__ void __launch_bounds__(kTraceBlock, SPC_PERSIST_MIN_BLOCKS)
k_trace_persist(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float4* __restrict__ rays,
                const int* __restrict__ n_dev, int mult, int64_t n_host, int cull_back, int fetch_threshold, int postpone_div,
                float4* __restrict__ hits, uint8_t* __restrict__ visible, unsigned long long* __restrict__ counter,
                unsigned long long* __restrict__ visit_counters = nullptr, const CamGen cam = CamGen()) {
    __shared__ uint2 s_stack[kSmStack * kTraceBlock];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);
    uint2 lstack[kLocStack + 2 * kMaxBvhDepth];   // per tree level: one node group, one parked and one postponed triangle group
    const int64_t n = n_dev ? (int64_t)__ldg(n_dev) * mult : n_host;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    Trav s;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(s_stack + threadIdx.x);
    const uint32_t lut = (uint32_t)__cvta_generic_to_shared(&s_lut);
    int64_t ray = -1;
    bool exhausted = false;
    unsigned cn = 0, ct = 0;
    while (true) {
        unsigned idle = __ballot_sync(0xffffffffu, ray < 0);
        if (idle) {
            if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= fetch_threshold)) {
                const int cnt = __popc(idle);
                const int leader = __ffs(idle) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(counter, (unsigned long long)cnt);
                base = __shfl_sync(0xffffffffu, base, leader);
                if (ray < 0) {
                    const int64_t my = (int64_t)base + __popc(idle & lt_mask);
                    if (my < n) {
                        float4 ro, rd;
                        if (RAYGEN) {
                            const unsigned pixel = cam.pixels ? (unsigned)__ldg(cam.pixels + my) : (unsigned)my;
                            uint32_t seed = tea<4>(pixel, cam.sample_index);
                            float jx = 0.5f, jy = 0.5f;
                            if (cam.sample_index != 0) {
                                jx = rnd(seed);
                                jy = rnd(seed);
                            }
                            const float3 d = camera_dir_exact(cam.U, cam.V, cam.W, pixel % cam.width, pixel / cam.width, cam.width, cam.height, jx, jy);
                            ro = make_float4(cam.eye.x, cam.eye.y, cam.eye.z, 1e-3f);   // SCENE_EPSILON
                            rd = make_float4(d.x, d.y, d.z, 1e16f);
                        } else {
                            ro = __ldg(rays + 2 * my);
                            rd = __ldg(rays + 2 * my + 1);
                        }
                        if (ANYHIT && !(rd.w > ro.w)) {
                            visible[my] = 1;   // empty interval: an unused connection slot
                        } else {
                            TravRay r{ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w, rd.w};
                            trav_init(s, r);
                            s.sbase = sbase;
                            ray = my;
                        }
                    }
                }
                if ((int64_t)base + cnt >= n) exhausted = true;
                idle = __ballot_sync(0xffffffffu, ray < 0);
            }
            if (exhausted && idle == 0xffffffffu) break;
        }
        if (ray >= 0) {
            if (trav_step2<ANYHIT, COUNT>(nodes, tris, s, cull_back != 0, kTraceBlock, lstack, lut, postpone_div, cn, ct)) {
                if (ANYHIT) {
                    visible[ray] = s.best_prim >= 0 ? 0 : 1;
                } else {
                    const bool hit = s.best_prim >= 0;
                    hits[ray] = make_float4(hit ? s.tcur : 0.0f, s.best_u, s.best_v, __int_as_float(s.best_prim));
                }
                ray = -1;
            }
        }
    }
    if (COUNT) {   // the instrumented twin: nodes visited / triangles tested by exactly this traversal order
        for (int o = 16; o > 0; o >>= 1) {
            cn += __shfl_xor_sync(0xffffffffu, cn, o);
            ct += __shfl_xor_sync(0xffffffffu, ct, o);
        }
        if (lane == 0) {
            atomicAdd(visit_counters + 0, (unsigned long long)cn);
            atomicAdd(visit_counters + 1, (unsigned long long)ct);
        }
    }
}

static int trace_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

struct Knobs {
    int max_blocks, fetch_t, postpone_div;
};
static const Knobs& persist_knobs() {
    static const Knobs knobs = {trace_env("SPC_TRACE_BLOCKS_PER_SM", 16), trace_env("SPC_FETCH_THRESHOLD", 6), trace_env("SPC_POSTPONE_DIV", 5)};
    return knobs;
}

template <bool ANYHIT>
static void launch_persist(Context& ctx, const spc_ray* rays, const int* n_dev, int mult, int64_t n_max, int flags, spc_hit* hits, uint8_t* visible,
                           unsigned long long* visit_counters = nullptr) {
    // process-wide tuning knobs (environment, read once) and the per-device occupancy of this kernel (queried once per context:
    // occupancy is a property of the device the context lives on)
    const Knobs& knobs = persist_knobs();
    int& cached = ctx.persist_blocks[ANYHIT ? 1 : 0];
    if (cached == 0) {
        int b = 0;
        SPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace_persist<ANYHIT>, kTraceBlock, 0));
        cached = std::max(1, std::min(b, knobs.max_blocks));
    }
    const int blocks_per_sm = cached, fetch_t = knobs.fetch_t, postpone_div = knobs.postpone_div;
    if (ctx.fetch_counters.n < 256) {
        ctx.fetch_counters.alloc(256);
        ctx.fetch_slot = 0;
    }
    unsigned long long* counter = ctx.fetch_counters.p + (ctx.fetch_slot++ & 255);
    SPC_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx.stream));
    int64_t blocks = (n_max + kTraceBlock - 1) / kTraceBlock;
    const int bps = ctx.trace_blocks_per_sm > 0 ? std::min(ctx.trace_blocks_per_sm, blocks_per_sm) : blocks_per_sm;
    const int64_t cap = (int64_t)ctx.sm_count * bps;
    if (blocks > cap) blocks = cap;
    const int cull = (flags & SPC_RAYFLAG_CULL_BACK_FACING) ? 1 : 0;
    if (visit_counters)
        k_trace_persist<ANYHIT, true><<<(unsigned)blocks, kTraceBlock, 0, ctx.stream>>>(ctx.bvh.nodes.p, ctx.bvh.tris.p, (const float4*)rays, n_dev, mult, n_max, cull,
                                                                                       fetch_t, postpone_div, (float4*)hits, visible, counter, visit_counters);
    else
        k_trace_persist<ANYHIT, false><<<(unsigned)blocks, kTraceBlock, 0, ctx.stream>>>(ctx.bvh.nodes.p, ctx.bvh.tris.p, (const float4*)rays, n_dev, mult, n_max, cull,
                                                                                        fetch_t, postpone_div, (float4*)hits, visible, counter, nullptr);
    SPC_CUDA(cudaGetLastError());
    ctx.launches++;
}

static bool use_persist() {
    static const int mode = trace_env("SPC_TRACE_PERSISTENT", 1);
    return mode != 0;
}

// Queue variants for the wavefront render passes: the number of rays is a device-side counter written by
// the previous stage (times `mult` rays per queue entry), the grid is fixed (a multiple of the SM count) and
// strides over the queue, so no host round trip is needed between bounces.
__global__ void __launch_bounds__(kTraceBlock)
k_trace_closest_q(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float4* __restrict__ rays,
                  const int* __restrict__ n_dev, int mult, int cull_back, float4* __restrict__ hits) {
    __shared__ uint2 s_stack[kSmStack * kTraceBlock];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);
    const int64_t n = (int64_t)__ldg(n_dev) * mult;
    unsigned cn = 0, ct = 0;
    for (int64_t i = (int64_t)blockIdx.x * kTraceBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kTraceBlock) {
        const float4 ro = __ldg(rays + 2 * i);
        const float4 rd = __ldg(rays + 2 * i + 1);
        TravRay r{ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w, rd.w};
        TravHit h;
        traverse_bvh8<false, false>(nodes, tris, r, cull_back != 0, s_stack + threadIdx.x, kTraceBlock, h, cn, ct, s_lut);
        hits[i] = make_float4(h.t, h.u, h.v, __int_as_float(h.prim));
    }
}

__global__ void __launch_bounds__(kTraceBlock)
k_trace_occlusion_q(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float4* __restrict__ rays,
                    const int* __restrict__ n_dev, int mult, uint8_t* __restrict__ visible) {
    __shared__ uint2 s_stack[kSmStack * kTraceBlock];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);
    const int64_t n = (int64_t)__ldg(n_dev) * mult;
    unsigned cn = 0, ct = 0;
    for (int64_t i = (int64_t)blockIdx.x * kTraceBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kTraceBlock) {
        const float4 ro = __ldg(rays + 2 * i);
        const float4 rd = __ldg(rays + 2 * i + 1);
        if (!(rd.w > ro.w)) {   // empty interval: an unused connection slot
            visible[i] = 1;
            continue;
        }
        TravRay r{ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w, rd.w};
        TravHit h;
        const bool blocked = traverse_bvh8<true, false>(nodes, tris, r, false, s_stack + threadIdx.x, kTraceBlock, h, cn, ct, s_lut);
        visible[i] = blocked ? 0 : 1;
    }
}

void launch_trace_closest_camera(Context& ctx, const CamGen& cam, int64_t n_pixels, int flags, spc_hit* hits) {
    if (n_pixels <= 0) return;
    int& cached = ctx.persist_blocks[0];
    if (cached == 0) {
        int b = 0;
        SPC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_trace_persist<false>, kTraceBlock, 0));
        cached = std::max(1, std::min(b, persist_knobs().max_blocks));
    }
    if (ctx.fetch_counters.n < 256) {
        ctx.fetch_counters.alloc(256);
        ctx.fetch_slot = 0;
    }
    unsigned long long* counter = ctx.fetch_counters.p + (ctx.fetch_slot++ & 255);
    SPC_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), ctx.stream));
    const int bps = ctx.trace_blocks_per_sm > 0 ? std::min(ctx.trace_blocks_per_sm, cached) : cached;
    const int64_t blocks = std::min<int64_t>((n_pixels + kTraceBlock - 1) / kTraceBlock, (int64_t)ctx.sm_count * bps);
    k_trace_persist<false, false, true><<<(unsigned)blocks, kTraceBlock, 0, ctx.stream>>>(
        ctx.bvh.nodes.p, ctx.bvh.tris.p, nullptr, nullptr, 1, n_pixels, (flags & SPC_RAYFLAG_CULL_BACK_FACING) ? 1 : 0, persist_knobs().fetch_t,
        persist_knobs().postpone_div, (float4*)hits, nullptr, counter, nullptr, cam);
    SPC_CUDA(cudaGetLastError());
    ctx.launches++;
}

void launch_trace_closest_q(Context& ctx, const spc_ray* rays, const int* n_dev, int mult, int64_t n_max, int flags, spc_hit* hits) {
    if (n_max <= 0) return;
    if (use_persist()) {
        launch_persist<false>(ctx, rays, n_dev, mult, n_max, flags, hits, nullptr);
        return;
    }
    int64_t blocks = (n_max + kTraceBlock - 1) / kTraceBlock;
    const int64_t cap = (int64_t)ctx.sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_trace_closest_q<<<(unsigned)blocks, kTraceBlock, 0, ctx.stream>>>(ctx.bvh.nodes.p, ctx.bvh.tris.p, (const float4*)rays, n_dev, mult,
                                                                      (flags & SPC_RAYFLAG_CULL_BACK_FACING) ? 1 : 0, (float4*)hits);
    SPC_CUDA(cudaGetLastError());
    ctx.launches++;
}

void launch_trace_occlusion_q(Context& ctx, const spc_ray* rays, const int* n_dev, int mult, int64_t n_max, uint8_t* visible) {
    if (n_max <= 0) return;
    if (use_persist()) {
        launch_persist<true>(ctx, rays, n_dev, mult, n_max, 0, nullptr, visible);
        return;
    }
    int64_t blocks = (n_max + kTraceBlock - 1) / kTraceBlock;
    const int64_t cap = (int64_t)ctx.sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_trace_occlusion_q<<<(unsigned)blocks, kTraceBlock, 0, ctx.stream>>>(ctx.bvh.nodes.p, ctx.bvh.tris.p, (const float4*)rays, n_dev, mult, visible);
    SPC_CUDA(cudaGetLastError());
    ctx.launches++;
}

void launch_trace_closest(Context& ctx, const spc_ray* rays, int64_t n, int flags, spc_hit* hits,
                          unsigned long long* counters) {
    if (n <= 0) return;
    const int64_t blocks = (n + kTraceBlock - 1) / kTraceBlock;
    SPC_REQUIRE(blocks < 0x7fffffffLL, SPC_ERR_INVALID, "ray batch too large: %lld", (long long)n);
    const int cull = (flags & SPC_RAYFLAG_CULL_BACK_FACING) ? 1 : 0;
    // counted variants: the production kernel with two visit counters (what THIS kernel fetches), or -- option "count_canonical" --
    // the plain one-ray-per-lane kernel: strict front-to-back, t-pruned order, the traversal SURVEY.md section 8d defines the
    // algorithmic bytes by (independent of the production kernel's scheduling, so visiting more nodes cannot raise the roofline)
    if (use_persist() && !(counters && ctx.opt[OPT_COUNT_CANONICAL])) {
        launch_persist<false>(ctx, rays, nullptr, 1, n, flags, hits, nullptr, counters);
        return;
    }
    if (counters)
        k_trace_closest<true><<<(unsigned)blocks, kTraceBlock, 0, ctx.stream>>>(
            ctx.bvh.nodes.p, ctx.bvh.tris.p, (const float4*)rays, n, cull, (float4*)hits, counters);
    else
        k_trace_closest<false><<<(unsigned)blocks, kTraceBlock, 0, ctx.stream>>>(
            ctx.bvh.nodes.p, ctx.bvh.tris.p, (const float4*)rays, n, cull, (float4*)hits, nullptr);
    SPC_CUDA(cudaGetLastError());
    ctx.launches++;
}

void launch_trace_occlusion(Context& ctx, const spc_ray* rays, int64_t n, uint8_t* visible,
                            unsigned long long* counters) {
    if (n <= 0) return;
    const int64_t blocks = (n + kTraceBlock - 1) / kTraceBlock;
    SPC_REQUIRE(blocks < 0x7fffffffLL, SPC_ERR_INVALID, "ray batch too large: %lld", (long long)n);
    if (use_persist() && !(counters && ctx.opt[OPT_COUNT_CANONICAL])) {
        launch_persist<true>(ctx, rays, nullptr, 1, n, 0, nullptr, visible, counters);
        return;
    }
    if (counters)
        k_trace_occlusion<true><<<(unsigned)blocks, kTraceBlock, 0, ctx.stream>>>(
            ctx.bvh.nodes.p, ctx.bvh.tris.p, (const float4*)rays, n, visible, counters);
    else
        k_trace_occlusion<false><<<(unsigned)blocks, kTraceBlock, 0, ctx.stream>>>(
            ctx.bvh.nodes.p, ctx.bvh.tris.p, (const float4*)rays, n, visible, nullptr);
    SPC_CUDA(cudaGetLastError());
    ctx.launches++;
}

}  // namespace spc
