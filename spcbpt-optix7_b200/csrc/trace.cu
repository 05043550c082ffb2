// trace.cu -- wavefront ray-batch kernels: closest hit and occlusion over the compressed 8-wide BVH.
// Stand in for the optixTrace calls of the reference (cuProg.h:384-487); one lane per ray, rays
// and hits are 128-bit coalesced loads/stores (32 B in, 16 B out), traversal stacks live in shared
// memory (kSmStack entries per lane) with a local-memory tail.
#include "traverse.cuh"

namespace spc {

constexpr int kTraceBlock = 128;

template <bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
k_trace_closest(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                const float4* __restrict__ rays, int64_t n, int cull_back, float4* __restrict__ hits,
                unsigned long long* __restrict__ counters) {
    __shared__ uint2 s_stack[kSmStack * kTraceBlock];
    const int64_t i = (int64_t)blockIdx.x * kTraceBlock + threadIdx.x;
    unsigned cn = 0, ct = 0;
    if (i < n) {
        const float4 ro = __ldg(rays + 2 * i);
        const float4 rd = __ldg(rays + 2 * i + 1);
        TravRay r{ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w, rd.w};
        TravHit h;
        traverse_bvh8<false, COUNT>(nodes, tris, r, cull_back != 0, s_stack + threadIdx.x, kTraceBlock, h, cn, ct);
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
    const int64_t i = (int64_t)blockIdx.x * kTraceBlock + threadIdx.x;
    unsigned cn = 0, ct = 0;
    if (i < n) {
        const float4 ro = __ldg(rays + 2 * i);
        const float4 rd = __ldg(rays + 2 * i + 1);
        TravRay r{ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w, rd.w};
        TravHit h;
        const bool blocked = traverse_bvh8<true, COUNT>(nodes, tris, r, false, s_stack + threadIdx.x, kTraceBlock, h, cn, ct);
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

// Queue variants for the wavefront render passes: the number of rays is a device-side counter written by
// the previous stage (times `mult` rays per queue entry), the grid is fixed (a multiple of the SM count) and
// strides over the queue, so no host round trip is needed between bounces.
__global__ void __launch_bounds__(kTraceBlock)
k_trace_closest_q(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float4* __restrict__ rays,
                  const int* __restrict__ n_dev, int mult, int cull_back, float4* __restrict__ hits) {
    __shared__ uint2 s_stack[kSmStack * kTraceBlock];
    const int64_t n = (int64_t)__ldg(n_dev) * mult;
    unsigned cn = 0, ct = 0;
    for (int64_t i = (int64_t)blockIdx.x * kTraceBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kTraceBlock) {
        const float4 ro = __ldg(rays + 2 * i);
        const float4 rd = __ldg(rays + 2 * i + 1);
        TravRay r{ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w, rd.w};
        TravHit h;
        traverse_bvh8<false, false>(nodes, tris, r, cull_back != 0, s_stack + threadIdx.x, kTraceBlock, h, cn, ct);
        hits[i] = make_float4(h.t, h.u, h.v, __int_as_float(h.prim));
    }
}

__global__ void __launch_bounds__(kTraceBlock)
k_trace_occlusion_q(const float4* __restrict__ nodes, const float4* __restrict__ tris, const float4* __restrict__ rays,
                    const int* __restrict__ n_dev, int mult, uint8_t* __restrict__ visible) {
    __shared__ uint2 s_stack[kSmStack * kTraceBlock];
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
        const bool blocked = traverse_bvh8<true, false>(nodes, tris, r, false, s_stack + threadIdx.x, kTraceBlock, h, cn, ct);
        visible[i] = blocked ? 0 : 1;
    }
}

void launch_trace_closest_q(Context& ctx, const spc_ray* rays, const int* n_dev, int mult, int64_t n_max, int flags, spc_hit* hits) {
    if (n_max <= 0) return;
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
