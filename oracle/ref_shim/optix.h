// Stub <optix.h> for building parts of the reference tree on the host with g++ (the OptiX SDK is
// not in this image).  Test infrastructure: declares only the names the
// reference's headers need; the device intrinsics are implemented by ref_host.cpp.
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
#include <math.h>
#include <stdlib.h>
// nvcc/MSVC give the reference global float overloads of min/max (CUDA: fminf/fmaxf semantics)
static inline float max(float a, float b) { return fmaxf(a, b); }
static inline float min(float a, float b) { return fminf(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline double min(double a, double b) { return fmin(a, b); }
static inline float max(float a, double b) { return fmaxf(a, (float)b); }
static inline float max(double a, float b) { return fmaxf((float)a, b); }
#include "optix_shim_device.h"
