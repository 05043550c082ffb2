// stub of the cmake-generated sampleConfig.h for building the reference's scene loader on the host (oracle/_ref/ref_loader)
#pragma once
#ifndef SAMPLES_DIR
#define SAMPLES_DIR "/root/reference/src"
#endif
#define CUDA_NVRTC_OPTIONS "-O3"
#include <sutil/vec_math.h>
