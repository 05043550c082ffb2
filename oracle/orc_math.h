// orc_math.h -- scalar fp32 helpers of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see README in
// this directory): nothing under oracle/ is linked into or called by the product library.
//
// The vector helpers restate the sutil/vec_math.h semantics the reference's device code relies on
// (normalize = v * (1/sqrt(dot)), sutil/vec_math.h `normalize`; lerp = a + t*(b-a); fmaxf(float3)).
// Compile with -ffp-contract=off: fused operations appear only where written (orc_fma).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct f3 {
    float x, y, z;
};
static inline f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
static inline f3 mk3(float s) { return f3{s, s, s}; }
static inline f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
static inline f3 operator*(f3 a, f3 b) { return f3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
static inline f3 operator*(float s, f3 a) { return f3{a.x * s, a.y * s, a.z * s}; }
static inline f3 operator/(f3 a, float s) {
    // sutil/vec_math.h: operator/(float3, float) multiplies by the reciprocal
    const float inv = 1.0f / s;
    return a * inv;
}
static inline f3 operator/(f3 a, f3 b) { return f3{a.x / b.x, a.y / b.y, a.z / b.z}; }
static inline f3& operator+=(f3& a, f3 b) { a = a + b; return a; }
static inline f3& operator*=(f3& a, f3 b) { a = a * b; return a; }
static inline f3& operator*=(f3& a, float s) { a = a * s; return a; }
static inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 cross(f3 a, f3 b) { return f3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline float length(f3 a) { return std::sqrt(dot(a, a)); }
static inline f3 normalize(f3 a) {
    const float inv = 1.0f / std::sqrt(dot(a, a));
    return a * inv;
}
static inline float fmax3(f3 a) { return std::fmax(std::fmax(a.x, a.y), a.z); }
static inline float sum3(f3 a) { return a.x + a.y + a.z; }
static inline float lerpf(float a, float b, float t) { return a + t * (b - a); }
static inline f3 lerp3(f3 a, f3 b, float t) { return a + t * (b - a); }
static inline float clampf(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }

// ---- contract arithmetic (bit-identical to csrc/traverse.cuh c_dot / c_cross) -------------------
static inline float orc_fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
static inline float c_dot(f3 a, f3 b) { return orc_fma(a.z, b.z, orc_fma(a.y, b.y, a.x * b.x)); }
static inline f3 c_cross(f3 a, f3 b) {
    return f3{orc_fma(a.y, b.z, -(a.z * b.y)), orc_fma(a.z, b.x, -(a.x * b.z)), orc_fma(a.x, b.y, -(a.y * b.x))};
}

// ---- contract transcendental functions -------------------------------------------------------------
// sin/cos/log/pow/exp of a float = the fp64 libm function rounded once to fp32.  CUDA's fp64 functions
// (<= 2 ulp) and glibc's (< 1 ulp) then agree bit-for-bit in fp32 except when the exact value lies within
// ~2 fp64-ulp of an fp32 rounding boundary (probability ~1e-8 per call), so the product's kernels
// (csrc/shade.cuh cm_*) and this oracle produce identical bits.  The reference itself uses
// --use_fast_math intrinsics (__sinf, __powf, ...), which no host build can reproduce; its host-shim build
// (oracle/ref_shim/ref_host.cpp) is given the same definitions.
static inline float cm_sinf(float x) { return (float)std::sin((double)x); }
static inline float cm_cosf(float x) { return (float)std::cos((double)x); }
static inline float cm_logf(float x) { return (float)std::log((double)x); }
static inline float cm_expf(float x) { return (float)std::exp((double)x); }
static inline float cm_powf(float x, float y) { return (float)std::pow((double)x, (double)y); }

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// ---- RNG: src/cuda/random.h:31-68 ---------------------------------------------------------------
// tea<N>: N rounds of the Tiny Encryption Algorithm over (val0,val1); returns v0.
static inline uint32_t tea(unsigned rounds, uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (unsigned n = 0; n < rounds; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
// lcg: state = 1664525*state + 1013904223; returns the low 24 bits (random.h:48-54)
static inline uint32_t lcg(uint32_t& s) {
    s = 1664525u * s + 1013904223u;
    return s & 0x00ffffffu;
}
// rnd: lcg / 2^24 as float in [0,1) (random.h:63-67)
static inline float rnd(uint32_t& s) { return (float)lcg(s) / (float)0x01000000; }

}  // namespace orc
