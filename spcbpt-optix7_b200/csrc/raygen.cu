// raygen.cu -- ray-batch producers: the pinhole camera of __raygen__SPCBPT / __raygen__pinhole
// (raygen.cu:321-344 in the reference) and the two derived ray sets of the traversal microbench
// (cosine-bounce and shadow rays from the primary hits).
#include "geom.cuh"

namespace spc {

// d = 2*((idx+jitter)/dims) - 1 ; dir = normalize(d.x*U + d.y*V + W)      (raygen.cu:338-343)
// jitter = (0.5,0.5) at subframe 0 else (rnd,rnd) drawn left to right       (raygen.cu:335-336)
__device__ __forceinline__ float3 camera_dir(float3 U, float3 V, float3 W, unsigned x, unsigned y, unsigned w, unsigned h,
                                             float jx, float jy) {
    return camera_dir_exact(U, V, W, x, y, w, h, jx, jy);
}

__global__ void k_camera_rays(float3 eye, float3 U, float3 V, float3 W, unsigned w, unsigned h, unsigned subframe,
                              float4* __restrict__ rays) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const unsigned x = i % w, y = i / w;
    uint32_t seed = tea<4>(i, subframe);
    float jx = 0.5f, jy = 0.5f;
    if (subframe != 0) {
        jx = rnd(seed);
        jy = rnd(seed);
    }
    const float3 d = camera_dir(U, V, W, x, y, w, h, jx, jy);
    rays[2 * (size_t)i] = make_float4(eye.x, eye.y, eye.z, 1e-3f);
    rays[2 * (size_t)i + 1] = make_float4(d.x, d.y, d.z, 1e16f);
}

// kind 1: cosine-hemisphere bounce from the hit point, seed tea<4>(i,1)
// kind 2: shadow ray from the hit point to a uniform point of light 0, seed tea<4>(i,2); [1e-3, len-1e-3]
__global__ void k_bench_rays(int kind, const float4* __restrict__ rays_in, const float4* __restrict__ hits_in, int64_t n,
                             const float4* __restrict__ tri_pos, const float2* __restrict__ tri_uv,
                             const spc_light* __restrict__ lights, float4* __restrict__ rays_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 ro = rays_in[2 * i], rd = rays_in[2 * i + 1];
    const float4 h = hits_in[i];
    const int prim = __float_as_int(h.w);
    if (prim < 0) {
        rays_out[2 * i] = make_float4(ro.x, ro.y, ro.z, 1e-3f);
        rays_out[2 * i + 1] = make_float4(rd.x, rd.y, rd.z, -1.0f);   // empty interval: always a miss
        return;
    }
    const LocalGeom g = local_geometry(tri_pos, tri_uv, prim, h.y, h.z);
    float3 N = g.Ng;
    const float3 din = f3(rd.x, rd.y, rd.z);
    if (dot(N, din) > 0.f) N = -N;
    if (kind == 1) {
        uint32_t seed = tea<4>((uint32_t)i, 1u);
        const float r1 = rnd(seed), r2 = rnd(seed);
        const Onb onb(N);
        const float3 d = onb.inverse_transform(cosine_sample_hemisphere(r1, r2));
        rays_out[2 * i] = make_float4(g.P.x, g.P.y, g.P.z, 1e-3f);
        rays_out[2 * i + 1] = make_float4(d.x, d.y, d.z, 1e16f);
    } else {
        uint32_t seed = tea<4>((uint32_t)i, 2u);
        const float r1 = rnd(seed), r2 = rnd(seed), r3 = 1.f - r1 - r2;
        const spc_light& L = lights[0];
        const float3 lp = ld3(L.u) * r1 + ld3(L.v) * r2 + ld3(L.corner) * r3;   // lightSample::ReverseSample, cuProg.h:576-580
        const float3 b = lp - g.P;
        const float len = length(b);
        const float3 d = b / len;
        rays_out[2 * i] = make_float4(g.P.x, g.P.y, g.P.z, 1e-3f);
        rays_out[2 * i + 1] = make_float4(d.x, d.y, d.z, len - 1e-3f);         // visibilityTest, cuProg.h:466-475
    }
}

}  // namespace spc

using spc::Context;

extern "C" {

// Pinhole primaries of one subframe: cam = {eye, U, V, W} as 12 floats (MyParams eye/U/V/W,
// whitted.h:75-78).  Replaces the raygen part of __raygen__SPCBPT / __raygen__pinhole (raygen.cu:321-344).
int spc_gen_camera_rays(spc_context* ctxp, const float* cam12, int width, int height, int subframe, spc_ray* rays_dev) {
    if (!ctxp || !cam12 || !rays_dev || width <= 0 || height <= 0) {
        spc::set_error("spc_gen_camera_rays: bad arguments");
        return SPC_ERR_INVALID;
    }
    Context& c = ctxp->c;
    try {
        cudaSetDevice(c.device);
        const unsigned n = (unsigned)width * (unsigned)height;
        spc::k_camera_rays<<<(n + 255) / 256, 256, 0, c.stream>>>(
            make_float3(cam12[0], cam12[1], cam12[2]), make_float3(cam12[3], cam12[4], cam12[5]),
            make_float3(cam12[6], cam12[7], cam12[8]), make_float3(cam12[9], cam12[10], cam12[11]), (unsigned)width,
            (unsigned)height, (unsigned)subframe, (float4*)rays_dev);
        SPC_CUDA(cudaGetLastError());
        c.launches++;
    } catch (const spc::CudaFailure& f) { return f.code; }
    return SPC_OK;
}

// Microbench ray sets B (kind 1, incoherent cosine bounce) and C (kind 2, shadow rays to light 0)
// derived from a primary batch and its hits.
int spc_gen_bench_rays(spc_context* ctxp, int kind, const spc_ray* rays_in_dev, const spc_hit* hits_in_dev, int64_t n,
                       spc_ray* rays_out_dev, void* reserved) {
    (void)reserved;
    if (!ctxp || !rays_in_dev || !hits_in_dev || !rays_out_dev || n < 0 || (kind != 1 && kind != 2)) {
        spc::set_error("spc_gen_bench_rays: bad arguments");
        return SPC_ERR_INVALID;
    }
    Context& c = ctxp->c;
    if (!c.has_scene || (kind == 2 && c.geom.n_lights < 1)) {
        spc::set_error("spc_gen_bench_rays: needs a scene (with a light for kind 2)");
        return SPC_ERR_NO_SCENE;
    }
    try {
        cudaSetDevice(c.device);
        if (n > 0) {
            spc::k_bench_rays<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(
                kind, (const float4*)rays_in_dev, (const float4*)hits_in_dev, n, c.geom.tri_pos.p, c.geom.tri_uv.p,
                c.geom.lights.p, (float4*)rays_out_dev);
            SPC_CUDA(cudaGetLastError());
            c.launches++;
        }
    } catch (const spc::CudaFailure& f) { return f.code; }
    return SPC_OK;
}

}  // extern "C"
