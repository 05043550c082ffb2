// Host implementations of the OptiX / CUDA device intrinsics used by raygen.cu, hit_program.cu,
// cuProg.h, rmis.h and src/cuda/*.h (list: SURVEY.md section 8c T1).  State is thread-local so
// that the reference's programs can run one "launch index" at a time on the host.
#pragma once
#include <cstring>
struct RefShimState {
    uint3 launch_index, launch_dims;
    unsigned int payload[4];
    const void* sbt_data;
    unsigned int prim_index;
    float2 bary;
    float ray_tmax;
    float3 ray_dir;
};
extern thread_local RefShimState g_shim;
void ref_shim_trace(float3 o, float3 d, float tmin, float tmax, unsigned int flags, unsigned int sbt_offset,
                    unsigned int* p0, unsigned int* p1);

static inline uint3 optixGetLaunchIndex() { return g_shim.launch_index; }
static inline uint3 optixGetLaunchDimensions() { return g_shim.launch_dims; }
static inline unsigned int optixGetPayload_0() { return g_shim.payload[0]; }
static inline unsigned int optixGetPayload_1() { return g_shim.payload[1]; }
static inline unsigned int optixGetPayload_2() { return g_shim.payload[2]; }
static inline unsigned int optixGetPayload_3() { return g_shim.payload[3]; }
static inline void optixSetPayload_0(unsigned int v) { g_shim.payload[0] = v; }
static inline void optixSetPayload_1(unsigned int v) { g_shim.payload[1] = v; }
static inline void optixSetPayload_2(unsigned int v) { g_shim.payload[2] = v; }
static inline void optixSetPayload_3(unsigned int v) { g_shim.payload[3] = v; }
static inline CUdeviceptr optixGetSbtDataPointer() { return (CUdeviceptr)g_shim.sbt_data; }
static inline unsigned int optixGetPrimitiveIndex() { return g_shim.prim_index; }
static inline float2 optixGetTriangleBarycentrics() { return g_shim.bary; }
static inline float optixGetRayTmax() { return g_shim.ray_tmax; }
static inline float3 optixGetWorldRayDirection() { return g_shim.ray_dir; }
static inline float3 optixTransformPointFromObjectToWorldSpace(float3 p) { return p; }   // identity instances, scene_shift.cpp:241,322
static inline float3 optixTransformNormalFromObjectToWorldSpace(float3 n) { return n; }
static inline void optixIgnoreIntersection() {}
static inline void optixTerminateRay() {}
// two-register payload (pointer) form: cuProg.h:395,420,445
static inline void optixTrace(OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax, float, OptixVisibilityMask,
                              unsigned int flags, unsigned int sbt_offset, unsigned int, unsigned int, unsigned int& p0, unsigned int& p1) {
    ref_shim_trace(o, d, tmin, tmax, flags, sbt_offset, &p0, &p1);
}
// one-register payload (occlusion) form: cuProg.h:470,518
static inline void optixTrace(OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax, float, OptixVisibilityMask,
                              unsigned int flags, unsigned int sbt_offset, unsigned int, unsigned int, unsigned int& p0) {
    ref_shim_trace(o, d, tmin, tmax, flags, sbt_offset, &p0, nullptr);
}
static inline unsigned int __float_as_uint(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned int u) { float f; memcpy(&f, &u, 4); return f; }
float4 ref_shim_tex2D(cudaTextureObject_t tex, float u, float v);
template <typename T> static inline T tex2D(cudaTextureObject_t tex, float u, float v) { return ref_shim_tex2D(tex, u, v); }
