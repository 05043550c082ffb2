// Minimal <optix.h> for compiling the reference's cuda_thrust/device_thrust.cu with nvcc (the OptiX SDK is not in
// this image): only the three typedefs its headers mention.  Test infrastructure.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>
#include <cmath>
#include <cstdio>
typedef unsigned long long CUdeviceptr;
typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;
