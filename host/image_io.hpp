// image_io.hpp -- images in and out of the host driver.
//
// In:  textures named by `albedoTex` in a .scene file, decoded to RGBA8 exactly as the reference hands them to
//      CUDA (scene_shift.cpp:36-56: stbi_load(name, &w, &h, &ch, STBI_rgb_alpha), 8 bit, 4 channels, first row = top).
//      Formats: baseline and progressive JPEG, PNG (what the shipped scene uses), binary PPM/PGM, and this repo's own raw
//      cache "<file>.rgba8" (magic "SPCRGBA8", int32 w, int32 h, w*h*4 bytes).
// Out: what replaces the GL display of the reference (sutil::GLDisplay / CUDAOutputBuffer, optixPathTracer.cpp:638-651):
//      the tone-mapped uchar4 frame buffer as a binary PPM and the float4 accumulation buffer as a PFM.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace spchost {

struct ImageRGBA8 {
    int width = 0, height = 0;
    std::vector<uint8_t> rgba;   // row-major, top row first
};

bool load_image_rgba8(const std::string& path, ImageRGBA8& out, std::string& err);
bool decode_jpeg_rgba8(const uint8_t* data, size_t size, ImageRGBA8& out, std::string& err);   // jpeg_decode.cpp
bool decode_png_rgba8(const uint8_t* data, size_t size, ImageRGBA8& out, std::string& err);    // png_decode.cpp
bool write_rgba8_cache(const std::string& path, const ImageRGBA8& img);

// Pixel (x, y) of the render lives at index y*W + x with y = 0 the BOTTOM row (raygen.cu:335-344 maps launch index y
// to d.y = 2(y+jitter)/H - 1 along +V, and the reference displays the buffer through GL, origin bottom-left).
// PPM is written top row first (flipped), PFM bottom row first (its native order, negative scale = little endian).
bool write_ppm_from_uchar4(const std::string& path, const uint32_t* frame, int width, int height);
bool write_pfm_from_float4(const std::string& path, const float* accum4, int width, int height);
// colour PFM ("PF", little or big endian) back in: rgb = width * height * 3 floats, bottom row first as in the file
bool read_pfm_rgb(const std::string& path, std::vector<float>& rgb, int& width, int& height, std::string& err);
// relMSE of an image against a reference image, the error metric of BASELINE.json / SURVEY.md section 8d:
// mean over pixels and channels of (I - R)^2 / (R^2 + 0.01); non-finite terms are skipped and counted in *skipped
double rel_mse(const std::vector<float>& img, const std::vector<float>& ref, size_t* skipped);
// 8-bit RGB PNG of the frame buffer (top row first; deflate by the system zlib, filter 0), for viewers that do not read PPM
bool write_png_from_uchar4(const std::string& path, const uint32_t* frame, int width, int height);

}  // namespace spchost
