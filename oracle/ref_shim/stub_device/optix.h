// Device-side stand-in for <optix.h> (the OptiX SDK is not in this image): lets the reference's OWN
// raygen.cu / hit_program.cu / cuProg.h / rmis.h compile unmodified with nvcc for sm_100a (--use_fast_math, as the reference builds
// them, src/CMakeLists.txt:214-215) so that they can run on the B200 as the GPU-side reference arm (oracle/ref_shim/ref_device.cu).
// TEST / BASELINE INFRASTRUCTURE: built only into oracle/_ref/ (git-ignored); never linked into the product.
//
// OptiX keeps the per-ray state (payload registers, hit attributes, SBT record) in its own registers; here it lives in shared
// memory, one slot per thread of a 1-D block of REF_SHIM_BLOCK threads.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>
#include <cmath>
#include <cstdio>
typedef unsigned long long CUdeviceptr;
typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
enum {
    OPTIX_RAY_FLAG_NONE = 0,
    OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1,
    OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 4,
    OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES = 16
};

#ifndef REF_SHIM_BLOCK
#define REF_SHIM_BLOCK 128
#endif

struct RefShimState {
    unsigned int payload[2];
    unsigned long long sbt_data;
    unsigned int prim_index;
    float ray_tmax;
    float2 bary;
    float3 ray_dir;
    unsigned int pad_;
};

#ifdef __CUDACC__
extern __shared__ RefShimState ref_shim_state[];     // REF_SHIM_BLOCK slots (dynamic shared memory of the wrapper kernels)
__constant__ uint3 ref_shim_dims;                    // launch dimensions of the current optixLaunch stand-in (one translation unit includes this)
#define REF_SHIM ref_shim_state[threadIdx.x]

__device__ void ref_shim_trace(float3 o, float3 d, float tmin, float tmax, unsigned int flags, unsigned int* p0, unsigned int* p1);

static __forceinline__ __device__ uint3 optixGetLaunchDimensions() { return ref_shim_dims; }
static __forceinline__ __device__ uint3 optixGetLaunchIndex() {
    const unsigned int i = blockIdx.x * REF_SHIM_BLOCK + threadIdx.x;
    return make_uint3(i % ref_shim_dims.x, i / ref_shim_dims.x, 0u);
}
static __forceinline__ __device__ unsigned int optixGetPayload_0() { return REF_SHIM.payload[0]; }
static __forceinline__ __device__ unsigned int optixGetPayload_1() { return REF_SHIM.payload[1]; }
static __forceinline__ __device__ unsigned int optixGetPayload_2() { return 0u; }
static __forceinline__ __device__ unsigned int optixGetPayload_3() { return 0u; }
static __forceinline__ __device__ void optixSetPayload_0(unsigned int v) { REF_SHIM.payload[0] = v; }
static __forceinline__ __device__ void optixSetPayload_1(unsigned int v) { REF_SHIM.payload[1] = v; }
static __forceinline__ __device__ void optixSetPayload_2(unsigned int) {}
static __forceinline__ __device__ void optixSetPayload_3(unsigned int) {}
static __forceinline__ __device__ CUdeviceptr optixGetSbtDataPointer() { return REF_SHIM.sbt_data; }
// the four intrinsics getLocalGeometry (SUTIL_HOSTDEVICE) calls must exist for the host pass too
static __forceinline__ __host__ __device__ unsigned int optixGetPrimitiveIndex() {
#ifdef __CUDA_ARCH__
    return REF_SHIM.prim_index;
#else
    return 0u;
#endif
}
static __forceinline__ __host__ __device__ float2 optixGetTriangleBarycentrics() {
#ifdef __CUDA_ARCH__
    return REF_SHIM.bary;
#else
    return make_float2(0.f, 0.f);
#endif
}
static __forceinline__ __host__ __device__ float3 optixTransformPointFromObjectToWorldSpace(float3 p) { return p; }    // identity instances, scene_shift.cpp:241,322
static __forceinline__ __host__ __device__ float3 optixTransformNormalFromObjectToWorldSpace(float3 n) { return n; }
static __forceinline__ __device__ float optixGetRayTmax() { return REF_SHIM.ray_tmax; }
static __forceinline__ __device__ float3 optixGetWorldRayDirection() { return REF_SHIM.ray_dir; }
static __forceinline__ __device__ void optixIgnoreIntersection() {}
static __forceinline__ __device__ void optixTerminateRay() {}
// two-register payload (pointer) form: cuProg.h:395,420,445
static __forceinline__ __device__ void optixTrace(OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax, float, OptixVisibilityMask,
                                                  unsigned int flags, unsigned int, unsigned int, unsigned int, unsigned int& p0, unsigned int& p1) {
    ref_shim_trace(o, d, tmin, tmax, flags, &p0, &p1);
}
// one-register payload (occlusion) form: cuProg.h:470,518
static __forceinline__ __device__ void optixTrace(OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax, float, OptixVisibilityMask,
                                                  unsigned int flags, unsigned int, unsigned int, unsigned int, unsigned int& p0) {
    ref_shim_trace(o, d, tmin, tmax, flags, &p0, nullptr);
}
#endif
