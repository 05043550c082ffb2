// geom.cuh -- hit-point reconstruction, RNG and small vector helpers shared by the shading kernels.
#pragma once
#include "common.cuh"

namespace spc {

// ---- float3 helpers (plain fp32; the compiler may contract these, parity-critical code below
//      uses explicit intrinsics instead) ------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 f3(float s) { return make_float3(s, s, s); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { const float inv = 1.0f / s; return a * inv; }  // sutil/vec_math.h:483-487
__device__ __forceinline__ float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ void operator+=(float3& a, float3 b) { a = a + b; }
__device__ __forceinline__ void operator*=(float3& a, float3 b) { a = a * b; }
__device__ __forceinline__ void operator*=(float3& a, float s) { a = a * s; }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float length(float3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ float3 normalize(float3 a) { const float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }  // sutil/vec_math.h:545-549
__device__ __forceinline__ float fmax3(float3 a) { return fmaxf(fmaxf(a.x, a.y), a.z); }
__device__ __forceinline__ float sum3(float3 a) { return a.x + a.y + a.z; }
__device__ __forceinline__ float lerpf(float a, float b, float t) { return a + t * (b - a); }
__device__ __forceinline__ float3 lerp3(float3 a, float3 b, float t) { return a + t * (b - a); }
__device__ __forceinline__ float3 ld3(const spc_float3& v) { return f3(v.x, v.y, v.z); }
__device__ __forceinline__ void st3(spc_float3& d, float3 v) { d.x = v.x; d.y = v.y; d.z = v.z; }

// ---- pinhole camera ray (raygen.cu:338-343): dir = normalize(d.x*U + d.y*V + W), d = 2*((idx+jitter)/dims) - 1.
// Written with explicit IEEE operations in the source's evaluation order: the exact flavour (-fmad=false) gets the very bits the
// plain expression gave, and the fast flavour (FMA contraction, approximate division: build.py) generates the SAME primary rays,
// so primary-hit primitive ids are identical in both (tests/test_fast_flavour_gpu.py).
__device__ __forceinline__ float3 pixel_dir_exact(float dx, float dy, float3 U, float3 V, float3 W) {
    const float x = __fadd_rn(__fadd_rn(__fmul_rn(U.x, dx), __fmul_rn(V.x, dy)), W.x);
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(U.y, dx), __fmul_rn(V.y, dy)), W.y);
    const float z = __fadd_rn(__fadd_rn(__fmul_rn(U.z, dx), __fmul_rn(V.z, dy)), W.z);
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))));
    return f3(__fmul_rn(x, inv), __fmul_rn(y, inv), __fmul_rn(z, inv));
}
__device__ __forceinline__ float3 camera_dir_exact(float3 U, float3 V, float3 W, unsigned x, unsigned y, unsigned w, unsigned h, float jx, float jy) {
    const float dx = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(__fadd_rn((float)x, jx), (float)w)), 1.0f);
    const float dy = __fsub_rn(__fmul_rn(2.0f, __fdiv_rn(__fadd_rn((float)y, jy), (float)h)), 1.0f);
    return pixel_dir_exact(dx, dy, U, V, W);
}

// ---- RNG: src/cuda/random.h:31-68 (TEA-N seed, LCG stream, 24-bit floats) ------------------------
template <unsigned N>
__host__ __device__ __forceinline__ uint32_t tea(uint32_t val0, uint32_t val1) {
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (unsigned n = 0; n < N; n++) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
__host__ __device__ __forceinline__ uint32_t lcg(uint32_t& s) {
    s = 1664525u * s + 1013904223u;
    return s & 0x00ffffffu;
}
__host__ __device__ __forceinline__ float rnd(uint32_t& s) { return (float)lcg(s) / (float)0x01000000; }

// ---- hit point (getLocalGeometry, src/cuda/LocalGeometry.h:59-176) --------------------------------
// Contract arithmetic (bit-identical to oracle orc_local_geometry): w = (1-u)-v,
// P = fma(v,P2, fma(u,P1, w*P0)); Ng = normalize(c_cross(P1-P0, P2-P0)) with normalize =
// v * (1/sqrt(c_dot(v,v))); UV likewise.  No vertex normals exist in the reference's scenes
// (scene_shift.cpp:234), so N = Ng.
struct LocalGeom {
    float3 P, Ng;
    float2 uv;
    int material, light, mesh;
};

__device__ __forceinline__ float gc_dot(float3 a, float3 b) { return __fmaf_rn(a.z, b.z, __fmaf_rn(a.y, b.y, __fmul_rn(a.x, b.x))); }
__device__ __forceinline__ float3 gc_cross(float3 a, float3 b) {
    return f3(__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)));
}

__device__ __forceinline__ LocalGeom local_geometry(const float4* __restrict__ tri_pos, const float2* __restrict__ tri_uv,
                                                    int prim, float u, float v) {
    const float4 a = __ldg(tri_pos + 3 * (size_t)prim), b = __ldg(tri_pos + 3 * (size_t)prim + 1), c = __ldg(tri_pos + 3 * (size_t)prim + 2);
    LocalGeom g;
    const float w = __fsub_rn(__fsub_rn(1.0f, u), v);
    g.P.x = __fmaf_rn(v, c.x, __fmaf_rn(u, b.x, __fmul_rn(w, a.x)));
    g.P.y = __fmaf_rn(v, c.y, __fmaf_rn(u, b.y, __fmul_rn(w, a.y)));
    g.P.z = __fmaf_rn(v, c.z, __fmaf_rn(u, b.z, __fmul_rn(w, a.z)));
    const float3 e1 = f3(__fsub_rn(b.x, a.x), __fsub_rn(b.y, a.y), __fsub_rn(b.z, a.z));
    const float3 e2 = f3(__fsub_rn(c.x, a.x), __fsub_rn(c.y, a.y), __fsub_rn(c.z, a.z));
    const float3 n = gc_cross(e1, e2);
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(gc_dot(n, n)));
    g.Ng = f3(__fmul_rn(n.x, inv), __fmul_rn(n.y, inv), __fmul_rn(n.z, inv));
    const float2 t0 = __ldg(tri_uv + 3 * (size_t)prim), t1 = __ldg(tri_uv + 3 * (size_t)prim + 1), t2 = __ldg(tri_uv + 3 * (size_t)prim + 2);
    g.uv.x = __fmaf_rn(v, t2.x, __fmaf_rn(u, t1.x, __fmul_rn(w, t0.x)));
    g.uv.y = __fmaf_rn(v, t2.y, __fmaf_rn(u, t1.y, __fmul_rn(w, t0.y)));
    g.material = __float_as_int(a.w);
    g.light = __float_as_int(b.w);
    g.mesh = __float_as_int(c.w);
    return g;
}

// ---- Onb + cosine hemisphere (cuProg.h:81-124) -------------------------------------------------------
struct Onb {
    float3 t, b, n;
    __device__ __forceinline__ explicit Onb(float3 normal) {
        n = normal;
        if (fabsf(n.x) > fabsf(n.z)) b = f3(-n.y, n.x, 0.f);
        else b = f3(0.f, -n.z, n.y);
        b = normalize(b);
        t = cross(b, n);
    }
    __device__ __forceinline__ float3 inverse_transform(float3 p) const { return p.x * t + p.y * b + p.z * n; }
};
__device__ __forceinline__ float3 cosine_sample_hemisphere(float u1, float u2) {
    const float r = sqrtf(u1);
    const float phi = 2.0f * 3.14159265358979323846f * u2;
    float3 p;
    p.x = r * cosf(phi);
    p.y = r * sinf(phi);
    p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
    return p;
}

}  // namespace spc
