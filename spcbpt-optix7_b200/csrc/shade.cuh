// shade.cuh -- device-side shading library of the SPCBPT render path: materials + textures, the
// Disney BSDF, emitter sampling, subspace classification, two-stage light-vertex sampling, recursive
// MIS and the connection estimator.  Each function names the reference code whose behaviour it
// reproduces (paths under src/OptiXPathTracer unless noted).
//
// Arithmetic policy (DESIGN.md "bit parity"): the library is compiled with -fmad=false -prec-div=true
// -prec-sqrt=true -ftz=false, every fp32 expression is written in the evaluation order of the
// reference's source, and transcendental functions go through cm_* (fp64 libm rounded once to fp32).
// The CPU oracle is written to the same policy, so whole frames compare bit-for-bit.
#pragma once
#include "geom.cuh"

namespace spc {

// ---- contract transcendental functions (see oracle/orc_math.h cm_*) ------------------------------------
// Two flavours of the library (build.py):
//   exact (libspcbpt_b200.so, the default and the one every parity test runs): fp64 libm rounded once to fp32, no contraction --
//         frames compare bit-for-bit with the oracle;
//   fast  (libspcbpt_b200_fast.so, -DSPC_FAST_MATH): what the reference's own build does (src/CMakeLists.txt:214-215
//         --use_fast_math): the hardware's fp32 special-function intrinsics, FMA contraction and approximate division / square
//         root in the shading code.  Traversal, primary rays and the training / binning code are identical in both flavours
//         (explicit IEEE intrinsics; exact flags): hit primitive ids and t are bit-equal, shading agrees to ~1e-6 relative per
//         operation (tests/test_fast_flavour_gpu.py states the tolerances).
#ifdef SPC_FAST_MATH
__device__ __forceinline__ float cm_sinf(float x) { return __sinf(x); }
__device__ __forceinline__ float cm_cosf(float x) { return __cosf(x); }
__device__ __forceinline__ float cm_logf(float x) { return __logf(x); }
__device__ __forceinline__ float cm_expf(float x) { return __expf(x); }
__device__ __forceinline__ float cm_powf(float x, float y) { return __powf(x, y); }
// the few fp64 operations the reference's source spells with double literals (cuProg.h:892, rmis.h:183, ...)
__device__ __forceinline__ float cm_div64(float a, float b) { return a / b; }
#else
__device__ __forceinline__ float cm_sinf(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cm_cosf(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float cm_logf(float x) { return (float)log((double)x); }
__device__ __forceinline__ float cm_expf(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float cm_powf(float x, float y) { return (float)pow((double)x, (double)y); }
__device__ __forceinline__ float cm_div64(float a, float b) { return (float)((double)a / (double)b); }
#endif

#define SPC_PI_F 3.14159265358979323846f
#define SPC_PI_D 3.14159265358979323846
#define SPC_SCENE_EPS 1e-3f   // SCENE_EPSILON, cuProg.h:39

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ float3 operator+(float3 a, float b) { return f3(a.x + b, a.y + b, a.z + b); }

// ---------------------------------------------------------------------------------------------
// what the kernels read: scene tables + the caller's MyParams + runtime constants
// ---------------------------------------------------------------------------------------------
struct DevScene {
    const float4*    tri_pos;
    const float2*    tri_uv;
    const spc_pbr*   materials;
    const spc_light* lights;
    const uint8_t*   tex_data;
    const int4*      tex_desc;
    const float4*    nodes;
    const float4*    tris;
    const float*     mat_log_cc;   // per material: cm_logf(a^2), a = lerp(0.1, 0.001, clearcoatGloss) (GTR1's only transcendental)
    int              n_lights;
    int              n_materials;
};

struct DevFrame {
    DevScene   sc;
    spc_params p;            // MyParams as given to spc_set_params (device pointers)
    int        K;            // NUM_SUBSPACE (optixPathTracer.h:31), runtime here
    int        connections;  // CONNECTION_N (:37)
    int        max_depth;    // literal 50 in raygen.cu:361,668 unless MyParams::max_depth > 0
    uint32_t   seed_offset;  // eye/pt seeds use tea<4>(pixel, subframe_index * seed_stride + seed_offset): (0, 1) = the reference's
    const short* lvc_xlabel; // eye-tree label of every LVC slot (k_lvc_xlabel), or null: walk the tree where it is needed
    const float4* eye_ctree;   // compact copies of subspace_info.eye_tree / light_tree (spc_tree_to_device) or null: walk the
    const float4* light_ctree; // reference-layout nodes
    const int* gamma_guide;  // guide tables of the CMFGamma rows ([K][K+1], guide.cu) or null: the reference's bisect
    const int* lvc_guide;    // guide tables of the per-subspace cmfs (table of subspace b at jump_bias + b, size + 1 entries) or null
    uint32_t   seed_stride;  // streams; other values when subframes are partitioned across GPUs or frame lanes (each keeps its own
                             // running mean over ITS subframes, numbered 0,1,2,... locally)
};

struct Pbr {   // the fields of MaterialData::Pbr the BSDF reads (src/cuda/MaterialData.h:78-97)
    float3 base_color;
    float  metallic, roughness, specular, specularTint, subsurface, sheen, sheenTint, clearcoat, clearcoatGloss;
    float  log_cc;   // cm_logf(a^2) of the clearcoat lobe, precomputed per material by the same device function (bit-identical)
    bool   brdf;
};

__device__ __forceinline__ Pbr load_pbr(const DevScene& sc, int id) {
    const float4* q = reinterpret_cast<const float4*>(sc.materials + id);   // spc_pbr is 16-aligned, 144 B
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    Pbr m;
    m.base_color = f3(a.x, a.y, a.z);
    m.metallic = b.x; m.roughness = b.y; m.specular = b.z; m.specularTint = b.w;
    m.subsurface = c.x; /* anisotropic c.y unused */ m.sheen = c.z; m.sheenTint = c.w;
    const float2 d = __ldg(reinterpret_cast<const float2*>(q + 3));
    m.clearcoat = d.x; m.clearcoatGloss = d.y;
    m.log_cc = __ldg(sc.mat_log_cc + id);
    m.brdf = sc.materials[id].brdf != 0;
    return m;
}

// fp32 bilinear fetch, wrap addressing, texel centres at +0.5 (cudaAddressModeWrap + cudaFilterModeLinear,
// scene_shift.cpp:57-60) with fp32 weights: CUDA's 9-bit filter weights are not reproducible on a host.
__device__ __forceinline__ float3 tex_fetch(const DevScene& sc, int tex, float u, float v) {
    const int4 d = __ldg(sc.tex_desc + tex);
    const int w = d.y, h = d.z;
    const uint8_t* px = sc.tex_data + d.x;
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    int x0 = (int)fx % w; if (x0 < 0) x0 += w;
    int x1 = ((int)fx + 1) % w; if (x1 < 0) x1 += w;
    int y0 = (int)fy % h; if (y0 < 0) y0 += h;
    int y1 = ((int)fy + 1) % h; if (y1 < 0) y1 += h;
    const uchar4 t00 = __ldg(reinterpret_cast<const uchar4*>(px) + (size_t)y0 * w + x0);
    const uchar4 t10 = __ldg(reinterpret_cast<const uchar4*>(px) + (size_t)y0 * w + x1);
    const uchar4 t01 = __ldg(reinterpret_cast<const uchar4*>(px) + (size_t)y1 * w + x0);
    const uchar4 t11 = __ldg(reinterpret_cast<const uchar4*>(px) + (size_t)y1 * w + x1);
    const float k = 1.0f / 255.0f;
    float3 out;
    {
        const float a = t00.x * k + ax * (t10.x * k - t00.x * k), b = t01.x * k + ax * (t11.x * k - t01.x * k);
        out.x = a + ay * (b - a);
    }
    {
        const float a = t00.y * k + ax * (t10.y * k - t00.y * k), b = t01.y * k + ax * (t11.y * k - t01.y * k);
        out.y = a + ay * (b - a);
    }
    {
        const float a = t00.z * k + ax * (t10.z * k - t00.z * k), b = t01.z * k + ax * (t11.z * k - t01.z * k);
        out.z = a + ay * (b - a);
    }
    return out;
}

// ColorTexSample (hit_program.cu:182-198) + sampleTexture (src/cuda/LocalShading.h:37-53) + linearize
// (cuProg.h:361-368).  RoughnessAndMetallicTexSample (hit_program.cu:199-209) multiplies by 1: no scene
// sets metallic_roughness_tex (scene_shift.cpp:64-91).
__device__ __forceinline__ Pbr shade_pbr(const DevScene& sc, int id, float2 uv) {
    Pbr m = load_pbr(sc, id);
    const spc_texture_ref& tr = sc.materials[id].base_color_tex;
    if (tr.tex != 0) {
        const float sx = uv.x * tr.texcoord_scale[0], sy = uv.y * tr.texcoord_scale[1];
        const float rx = tr.texcoord_rotation[0], ry = tr.texcoord_rotation[1];
        const float tu = (sx * ry + sy * rx) + tr.texcoord_offset[0];
        const float tv = (sx * (-rx) + sy * ry) + tr.texcoord_offset[1];
        const float3 c = tex_fetch(sc, (int)tr.tex - 1, tu, tv);
        m.base_color = f3(cm_powf(c.x, 2.2f), cm_powf(c.y, 2.2f), cm_powf(c.z, 2.2f));
    }
    m.roughness *= 1.0f;
    m.metallic *= 1.0f;
    return m;
}

// getLocalGeometry (src/cuda/LocalGeometry.h:59-176) with identity instance transforms and no vertex normals
// (scene_shift.cpp:234,241): P = (1-u-v) P0 + u P1 + v P2, Ng = normalize(cross(P1-P0, P2-P0)), UV likewise --
// plain fp32 in the source's evaluation order (the traversal's fused contract arithmetic is NOT used here).
__device__ __forceinline__ LocalGeom hit_geometry(const DevScene& sc, int prim, float bu, float bv) {
    const float4 a = __ldg(sc.tri_pos + 3 * (size_t)prim), b = __ldg(sc.tri_pos + 3 * (size_t)prim + 1), c = __ldg(sc.tri_pos + 3 * (size_t)prim + 2);
    const float3 v0 = f3(a.x, a.y, a.z), v1 = f3(b.x, b.y, b.z), v2 = f3(c.x, c.y, c.z);
    // explicit IEEE operations in the source's order (= what the plain expressions give without contraction): the vertex position and
    // normal are the inputs of the subspace classification, so they are kept bit-identical in the exact and the fast flavour
    LocalGeom g;
    const float w = __fsub_rn(__fsub_rn(1.0f, bu), bv);
    g.P.x = __fadd_rn(__fadd_rn(__fmul_rn(v0.x, w), __fmul_rn(v1.x, bu)), __fmul_rn(v2.x, bv));
    g.P.y = __fadd_rn(__fadd_rn(__fmul_rn(v0.y, w), __fmul_rn(v1.y, bu)), __fmul_rn(v2.y, bv));
    g.P.z = __fadd_rn(__fadd_rn(__fmul_rn(v0.z, w), __fmul_rn(v1.z, bu)), __fmul_rn(v2.z, bv));
    const float3 e1 = f3(__fsub_rn(v1.x, v0.x), __fsub_rn(v1.y, v0.y), __fsub_rn(v1.z, v0.z));
    const float3 e2 = f3(__fsub_rn(v2.x, v0.x), __fsub_rn(v2.y, v0.y), __fsub_rn(v2.z, v0.z));
    const float3 n = f3(__fsub_rn(__fmul_rn(e1.y, e2.z), __fmul_rn(e1.z, e2.y)), __fsub_rn(__fmul_rn(e1.z, e2.x), __fmul_rn(e1.x, e2.z)),
                        __fsub_rn(__fmul_rn(e1.x, e2.y), __fmul_rn(e1.y, e2.x)));
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(n.x, n.x), __fmul_rn(n.y, n.y)), __fmul_rn(n.z, n.z))));
    g.Ng = f3(__fmul_rn(n.x, inv), __fmul_rn(n.y, inv), __fmul_rn(n.z, inv));
    const float2 t0 = __ldg(sc.tri_uv + 3 * (size_t)prim), t1 = __ldg(sc.tri_uv + 3 * (size_t)prim + 1), t2 = __ldg(sc.tri_uv + 3 * (size_t)prim + 2);
    g.uv.x = __fadd_rn(__fadd_rn(__fmul_rn(w, t0.x), __fmul_rn(bu, t1.x)), __fmul_rn(bv, t2.x));
    g.uv.y = __fadd_rn(__fadd_rn(__fmul_rn(w, t0.y), __fmul_rn(bu, t1.y)), __fmul_rn(bv, t2.y));
    g.material = __float_as_int(a.w);
    g.light = __float_as_int(b.w);
    g.mesh = __float_as_int(c.w);
    return g;
}

// ---------------------------------------------------------------------------------------------
// Disney BSDF: Tracer::Eval / Sample / Pdf (cuProg.h:684-899)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sqr(float x) { return x * x; }
__device__ __forceinline__ float SchlickFresnel(float u) {
    const float m = clampf(1.0f - u, 0.0f, 1.0f);
    const float m2 = m * m;
    return m2 * m2 * m;
}
// GTR1 (cuProg.h:693-699); log_a2 = cm_logf(a*a), taken from the per-material table (k_material_tables) instead of
// being re-evaluated in fp64 ten times per connection
__device__ __forceinline__ float GTR1(float NDotH, float a, float log_a2) {
    if (a >= 1.0f) return (1.0f / SPC_PI_F);
    const float a2 = a * a;
    const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return (a2 - 1.0f) / (SPC_PI_F * log_a2 * t);
}
__device__ __forceinline__ float GTR2(float NDotH, float a) {
    const float a2 = a * a;
    const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return a2 / (SPC_PI_F * t * t);
}
__device__ __forceinline__ float smithG_GGX(float NDotv, float alphaG) {
    const float a = alphaG * alphaG;
    const float b = NDotv * NDotv;
    return 1.0f / (NDotv + sqrtf(a + b - a * b));
}

static __device__ __noinline__ float3 bsdf_eval(const Pbr& mat, float3 N, float3 V, float3 L) {
    const float NDotL = dot(N, L);
    const float NDotV = dot(N, V);
    if (NDotL <= 0.0f || NDotV <= 0.0f) return f3(0.0f);
    const float3 H = normalize(L + V);
    const float NDotH = dot(N, H);
    const float LDotH = dot(L, H);
    const float3 Cdlin = mat.base_color;
    const float Cdlum = 0.3f * Cdlin.x + 0.6f * Cdlin.y + 0.1f * Cdlin.z;
    const float3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : f3(1.0f);
    const float3 Cspec0 = lerp3(mat.specular * 0.08f * lerp3(f3(1.0f), Ctint, mat.specularTint), Cdlin, mat.metallic);
    const float3 Csheen = lerp3(f3(1.0f), Ctint, mat.sheenTint);
    const float FL = SchlickFresnel(NDotL), FV = SchlickFresnel(NDotV);
    const float Fd90 = 0.5f + 2.0f * LDotH * LDotH * mat.roughness;
    const float Fd = lerpf(1.0f, Fd90, FL) * lerpf(1.0f, Fd90, FV);
    const float Fss90 = LDotH * LDotH * mat.roughness;
    const float Fss = lerpf(1.0f, Fss90, FL) * lerpf(1.0f, Fss90, FV);
    const float ss = 1.25f * (Fss * (1.0f / (NDotL + NDotV) - 0.5f) + 0.5f);
    const float a = fmaxf(0.001f, mat.roughness);
    const float Ds = GTR2(NDotH, a);
    const float FH = SchlickFresnel(LDotH);
    const float3 Fs = lerp3(Cspec0, f3(1.0f), FH);
    const float roughg = sqr(mat.roughness * 0.5f + 0.5f);
    const float Gs = smithG_GGX(NDotL, roughg) * smithG_GGX(NDotV, roughg);
    const float3 Fsheen = FH * mat.sheen * Csheen;
    const float Dr = GTR1(NDotH, lerpf(0.1f, 0.001f, mat.clearcoatGloss), mat.log_cc);
    const float Fr = lerpf(0.04f, 1.0f, FH);
    const float Gr = smithG_GGX(NDotL, 0.25f) * smithG_GGX(NDotV, 0.25f);
    const float3 out = ((1.0f / SPC_PI_F) * lerpf(Fd, ss, mat.subsurface) * Cdlin + Fsheen) * (1.0f - mat.metallic) + Gs * Fs * Ds +
                       0.25f * mat.clearcoat * Gr * Fr * Dr;
    return out;
}

__device__ __forceinline__ float3 bsdf_sample(const Pbr& mat, float3 N, float3 V, uint32_t& seed) {
    float3 dir;
    const float probability = rnd(seed);
    const float diffuseRatio = 0.5f * (1.0f - mat.metallic);
    const float r1 = rnd(seed);
    const float r2 = rnd(seed);
    const Onb onb(N);
    if (probability < diffuseRatio) {
        const float r = sqrtf(r1);
        const float phi = 2.0f * SPC_PI_F * r2;
        float3 p;
        p.x = r * cm_cosf(phi);
        p.y = r * cm_sinf(phi);
        p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
        dir = onb.inverse_transform(p);
    } else {
        const float a = fmaxf(0.001f, mat.roughness);
        const float phi = r1 * 2.0f * SPC_PI_F;
        const float cosTheta = sqrtf((1.0f - r2) / (1.0f + (a * a - 1.0f) * r2));
        const float sinTheta = sqrtf(1.0f - (cosTheta * cosTheta));
        const float sinPhi = cm_sinf(phi);
        const float cosPhi = cm_cosf(phi);
        float3 half = f3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
        half = onb.inverse_transform(half);
        dir = 2.0f * dot(V, half) * half - V;
    }
    return dir;
}

static __device__ __noinline__ float bsdf_pdf(const Pbr& mat, float3 n, float3 V, float3 L) {
    const float specularAlpha = fmaxf(0.001f, mat.roughness);
    const float clearcoatAlpha = lerpf(0.1f, 0.001f, mat.clearcoatGloss);
    const float diffuseRatio = 0.5f * (1.f - mat.metallic);
    const float specularRatio = 1.f - diffuseRatio;
    const float3 half = normalize(L + V);
    const float cosTheta = fabsf(dot(half, n));
    const float pdfGTR2 = GTR2(cosTheta, specularAlpha) * cosTheta;
    const float pdfGTR1 = GTR1(cosTheta, clearcoatAlpha, mat.log_cc) * cosTheta;
    const float ratio = 1.0f / (1.0f + mat.clearcoat);
    // `/ (4.0 * abs(...))`: the literal is a double, so this one division is fp64 in the reference (cuProg.h:892)
#ifdef SPC_FAST_MATH
    const float pdfSpec = lerpf(pdfGTR1, pdfGTR2, ratio) / (4.0f * fabsf(dot(L, half)));
#else
    const float pdfSpec = (float)((double)lerpf(pdfGTR1, pdfGTR2, ratio) / (4.0 * (double)fabsf(dot(L, half))));
#endif
    const float pdfDiff = fabsf(dot(L, n)) * (1.0f / SPC_PI_F);
    return diffuseRatio * pdfDiff + specularRatio * pdfSpec;
}

// ---------------------------------------------------------------------------------------------
// classification: classTree::tree_index (decisionTree/classTree_common.h:39-51) via labelUnit::getLabel
// (cuProg.h:1109-1123).  Null tree -> 0; DIR_JUDGE 0 -> a type-2 (direction) node compares against 0.
// One 56-byte node = mid (12) + 8 children (32) + label, type, leaf: only the child actually taken is read.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int tree_label(const spc_tree_node* __restrict__ root, float3 position, float3 normal) {
    if (!root) return 0;
    int node = 0;
    while (true) {
        const spc_tree_node* nd = root + node;
        if (__ldg(&nd->leaf)) return __ldg(&nd->label);
        const int type = __ldg(&nd->type);
        const float3 q = type == 0 ? position : (type == 1 ? normal : f3(0.0f));
        const float mx = __ldg(&nd->mid.x), my = __ldg(&nd->mid.y), mz = __ldg(&nd->mid.z);
        int ind = 0;
        ind += q.x > mx ? 1 : 0;
        ind += q.y > my ? 2 : 0;
        ind += q.z > mz ? 4 : 0;
        node = __ldg(&nd->child[ind]);
    }
}

// Compact trees.  The reference-layout node (56 B: mid, 8 children, label, type, leaf) costs six scalar loads per level, each its own
// L1 wavefront per lane once the lanes of a warp have diverged.  spc_tree_to_device keeps a second copy of 48 B = 3 x float4 per node:
// {mid.xyz, type} and the 8 children, where a leaf child is stored in its parent as 0x80000000 | label: one 128-bit load plus one
// 32-bit load per level, and no load at all for the leaf.  Same comparisons, same label.
__device__ __forceinline__ int ctree_step(const float4* __restrict__ t, int node, float3 position, float3 normal) {
    const float4 q = __ldg(t + (size_t)node * 3);
    const uint32_t type = __float_as_uint(q.w);
    const float3 p = type == 0 ? position : (type == 1 ? normal : f3(0.0f));
    int ind = 0;
    ind += p.x > q.x ? 1 : 0;
    ind += p.y > q.y ? 2 : 0;
    ind += p.z > q.z ? 4 : 0;
    return __ldg(reinterpret_cast<const int*>(t + (size_t)node * 3 + 1) + ind);
}
__device__ __forceinline__ int ctree_label(const float4* __restrict__ t, float3 position, float3 normal) {
    const uint32_t root = __float_as_uint(__ldg(t).w);
    if (root & 0x80000000u) return (int)(root & 0x7fffffffu);
    int node = 0;
    while (node >= 0) node = ctree_step(t, node, position, normal);
    return node & 0x7fffffff;
}
// both trees at once: two independent chains of dependent loads in flight
__device__ __forceinline__ void ctree_label2(const float4* __restrict__ ta, const float4* __restrict__ tb, float3 position, float3 normal,
                                             int& labelA, int& labelB) {
    const uint32_t ra = __float_as_uint(__ldg(ta).w), rb = __float_as_uint(__ldg(tb).w);
    int na = (ra & 0x80000000u) ? (int)ra : 0, nb = (rb & 0x80000000u) ? (int)rb : 0;
    while (na >= 0 || nb >= 0) {
        if (na >= 0) na = ctree_step(ta, na, position, normal);
        if (nb >= 0) nb = ctree_step(tb, nb, position, normal);
    }
    labelA = na & 0x7fffffff;
    labelB = nb & 0x7fffffff;
}
__device__ __forceinline__ int eye_tree_label(const DevFrame& fr, float3 position, float3 normal) {
    return fr.eye_ctree ? ctree_label(fr.eye_ctree, position, normal) : tree_label(fr.p.subspace_info.eye_tree, position, normal);
}
__device__ __forceinline__ int light_tree_label(const DevFrame& fr, float3 position, float3 normal) {
    return fr.light_ctree ? ctree_label(fr.light_ctree, position, normal) : tree_label(fr.p.subspace_info.light_tree, position, normal);
}

// the same walk down two trees at once (eye-tree label and light-tree label of one vertex): two independent chains of dependent
// loads in flight instead of one after the other
__device__ __forceinline__ void tree_label2(const spc_tree_node* __restrict__ rootA, const spc_tree_node* __restrict__ rootB, float3 position,
                                            float3 normal, int& labelA, int& labelB) {
    int nodeA = 0, nodeB = 0;
    bool doneA = rootA == nullptr, doneB = rootB == nullptr;
    labelA = 0;
    labelB = 0;
    while (!(doneA && doneB)) {
        const spc_tree_node* a = rootA + nodeA;
        const spc_tree_node* b = rootB + nodeB;
        int leafA = 1, leafB = 1, typeA = 0, typeB = 0;
        if (!doneA) { leafA = __ldg(&a->leaf); typeA = __ldg(&a->type); }
        if (!doneB) { leafB = __ldg(&b->leaf); typeB = __ldg(&b->type); }
        if (!doneA) {
            if (leafA) { labelA = __ldg(&a->label); doneA = true; }
            else {
                const float3 q = typeA == 0 ? position : (typeA == 1 ? normal : f3(0.0f));
                int ind = 0;
                ind += q.x > __ldg(&a->mid.x) ? 1 : 0;
                ind += q.y > __ldg(&a->mid.y) ? 2 : 0;
                ind += q.z > __ldg(&a->mid.z) ? 4 : 0;
                nodeA = __ldg(&a->child[ind]);
            }
        }
        if (!doneB) {
            if (leafB) { labelB = __ldg(&b->label); doneB = true; }
            else {
                const float3 q = typeB == 0 ? position : (typeB == 1 ? normal : f3(0.0f));
                int ind = 0;
                ind += q.x > __ldg(&b->mid.x) ? 1 : 0;
                ind += q.y > __ldg(&b->mid.y) ? 2 : 0;
                ind += q.z > __ldg(&b->mid.z) ? 4 : 0;
                nodeB = __ldg(&b->child[ind]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// emitter sampling: Tracer::lightSample (cuProg.h:554-666), QUAD lights
// ---------------------------------------------------------------------------------------------
struct LightSample {
    float3 position, emission, direction, normal;
    float  uvx, uvy, pdf, dir_pdf;
    int    subspaceId, light_id;
};
__device__ __forceinline__ void light_reverse_sample(const DevFrame& fr, int li, float r1, float r2, LightSample& s) {   // cuProg.h:571-601
    const spc_light& L = fr.sc.lights[li];
    const float r3 = 1 - r1 - r2;
    s.position = ld3(L.u) * r1 + ld3(L.v) * r2 + ld3(L.corner) * r3;
    s.emission = ld3(L.emission);
    s.normal = ld3(L.normal);
    s.pdf = cm_div64(1.0f, L.area);
    s.pdf /= (float)(unsigned)fr.sc.n_lights;
    s.uvx = r1; s.uvy = r2;
    const int xb = max(0, min((int)floorf(s.uvx * L.divLevel), L.divLevel - 1));
    const int yb = max(0, min((int)floorf(s.uvy * L.divLevel), L.divLevel - 1));
    const int lightSpaceId = L.ssBase + xb * L.divLevel + yb;
    s.subspaceId = fr.K - lightSpaceId - 1;
    s.light_id = L.id;
}
__device__ __forceinline__ int pick_light(const DevFrame& fr, uint32_t& seed) {   // raygen.cu:639, cuProg.h:624
    const int n = fr.sc.n_lights;
    return max(0, min((int)floorf(rnd(seed) * (float)(unsigned)n), n - 1));
}
__device__ __forceinline__ void light_trace_mode(LightSample& s, uint32_t& seed) {   // cuProg.h:648-665
    const float r1 = rnd(seed);
    const float r2 = rnd(seed);
    const Onb onb(s.normal);
    const float r = sqrtf(r1);
    const float phi = 2.0f * SPC_PI_F * r2;
    float3 p;
    p.x = r * cm_cosf(phi);
    p.y = r * cm_sinf(phi);
    p.z = sqrtf(fmaxf(0.0f, 1.0f - p.x * p.x - p.y * p.y));
    s.direction = onb.inverse_transform(p);
    s.dir_pdf = fabsf(dot(s.direction, s.normal)) / SPC_PI_F;
}

// ---------------------------------------------------------------------------------------------
// vertex in registers (same fields as BDPTVertex, BDPTVertex.h:9-70)
// ---------------------------------------------------------------------------------------------
struct Vtx {
    float3 position, normal, flux, color, lastPosition, RMIS_pointer_3;
    float2 uv;
    float  RMIS_pointer, last_lum, lastNormalProjection, pdf, singlePdf, lastSinglePdf;
    short  materialId, subspaceId, depth, lastZoneId, type;
    unsigned char isOrigin, inBrdf, lastBrdf, isBrdf, isLastVertex_direction;
};
__device__ __forceinline__ void vtx_zero(Vtx& v) {
    v.position = v.normal = v.flux = v.color = v.lastPosition = v.RMIS_pointer_3 = f3(0.f);
    v.uv = make_float2(0.f, 0.f);
    v.RMIS_pointer = v.last_lum = v.lastNormalProjection = v.pdf = v.singlePdf = v.lastSinglePdf = 0.f;
    v.materialId = v.subspaceId = v.depth = v.lastZoneId = 0;
    v.type = SPC_VTYPE_QUAD;   // BDPTVertex default (BDPTVertex.h:55)
    v.isOrigin = v.inBrdf = v.lastBrdf = v.isBrdf = v.isLastVertex_direction = 0;
}
// 120-byte records are 8-aligned: 15 x 64-bit accesses
__device__ __forceinline__ Vtx vtx_load(const spc_vertex* p) {
    const uint2* q = reinterpret_cast<const uint2*>(p);
    uint2 w[15];
#pragma unroll
    for (int i = 0; i < 15; i++) w[i] = q[i];
    Vtx v;
#define F(u) __uint_as_float(u)
    v.position = f3(F(w[0].x), F(w[0].y), F(w[1].x));
    v.normal = f3(F(w[1].y), F(w[2].x), F(w[2].y));
    v.flux = f3(F(w[3].x), F(w[3].y), F(w[4].x));
    v.color = f3(F(w[4].y), F(w[5].x), F(w[5].y));
    v.lastPosition = f3(F(w[6].x), F(w[6].y), F(w[7].x));
    v.RMIS_pointer_3 = f3(F(w[7].y), F(w[8].x), F(w[8].y));
    v.uv = make_float2(F(w[9].x), F(w[9].y));
    v.RMIS_pointer = F(w[10].x); v.last_lum = F(w[10].y);
    v.lastNormalProjection = F(w[11].x); v.pdf = F(w[11].y);
    v.singlePdf = F(w[12].x); v.lastSinglePdf = F(w[12].y);
#undef F
    v.materialId = (short)(w[13].x & 0xffffu); v.subspaceId = (short)(w[13].x >> 16);
    v.depth = (short)(w[13].y & 0xffffu); v.lastZoneId = (short)(w[13].y >> 16);
    v.type = (short)(w[14].x & 0xffffu);
    v.isOrigin = (unsigned char)((w[14].x >> 16) & 0xffu); v.inBrdf = (unsigned char)(w[14].x >> 24);
    v.lastBrdf = (unsigned char)(w[14].y & 0xffu); v.isBrdf = (unsigned char)((w[14].y >> 8) & 0xffu);
    v.isLastVertex_direction = (unsigned char)((w[14].y >> 16) & 0xffu);
    return v;
}
__device__ __forceinline__ void vtx_store(spc_vertex* p, const Vtx& v) {
    uint2* q = reinterpret_cast<uint2*>(p);
#define U(f) __float_as_uint(f)
    q[0] = make_uint2(U(v.position.x), U(v.position.y));
    q[1] = make_uint2(U(v.position.z), U(v.normal.x));
    q[2] = make_uint2(U(v.normal.y), U(v.normal.z));
    q[3] = make_uint2(U(v.flux.x), U(v.flux.y));
    q[4] = make_uint2(U(v.flux.z), U(v.color.x));
    q[5] = make_uint2(U(v.color.y), U(v.color.z));
    q[6] = make_uint2(U(v.lastPosition.x), U(v.lastPosition.y));
    q[7] = make_uint2(U(v.lastPosition.z), U(v.RMIS_pointer_3.x));
    q[8] = make_uint2(U(v.RMIS_pointer_3.y), U(v.RMIS_pointer_3.z));
    q[9] = make_uint2(U(v.uv.x), U(v.uv.y));
    q[10] = make_uint2(U(v.RMIS_pointer), U(v.last_lum));
    q[11] = make_uint2(U(v.lastNormalProjection), U(v.pdf));
    q[12] = make_uint2(U(v.singlePdf), U(v.lastSinglePdf));
#undef U
    q[13] = make_uint2((uint32_t)(uint16_t)v.materialId | ((uint32_t)(uint16_t)v.subspaceId << 16),
                       (uint32_t)(uint16_t)v.depth | ((uint32_t)(uint16_t)v.lastZoneId << 16));
    q[14] = make_uint2((uint32_t)(uint16_t)v.type | ((uint32_t)v.isOrigin << 16) | ((uint32_t)v.inBrdf << 24),
                       (uint32_t)v.lastBrdf | ((uint32_t)v.isBrdf << 8) | ((uint32_t)v.isLastVertex_direction << 16));
}

__device__ __forceinline__ void init_vertex_from_light_sample(const LightSample& s, Vtx& v) {   // raygen.cu:172-195
    v.position = s.position;
    v.normal = s.normal;
    v.flux = s.emission;
    v.pdf = s.pdf;
    v.singlePdf = v.pdf;
    v.isOrigin = 1;
    v.isBrdf = 0;
    v.subspaceId = (short)s.subspaceId;
    v.depth = 0;
    v.materialId = (short)s.light_id;
    v.RMIS_pointer = 1;
    v.uv = make_float2(s.uvx, s.uvy);
    v.type = SPC_VTYPE_QUAD;
}

// ---------------------------------------------------------------------------------------------
// recursive MIS (rmis.h)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ Pbr vertex_mat(const DevFrame& fr, const Vtx& v) {   // rmis::getMat, rmis.h:16-21
    Pbr m = load_pbr(fr.sc, v.materialId);
    m.base_color = v.color;
    return m;
}
__device__ __forceinline__ float Gamma(const DevFrame& fr, int eye_id, int light_id) {   // optixPathTracer.h:173-181
    const float* C = fr.p.subspace_info.CMFGamma;
    if (C && fr.p.subspace_info.Q) {
        const size_t i = (size_t)eye_id * fr.K + light_id;
        return light_id == 0 ? __ldg(C + i) : __ldg(C + i) - __ldg(C + i - 1);
    }
    return 1;
}
__device__ __forceinline__ float gamma_ss(const DevFrame& fr, int eye_id, int light_id) {   // optixPathTracer.h:182-189
    if (fr.p.subspace_info.CMFGamma && fr.p.subspace_info.Q) return Gamma(fr, eye_id, light_id) / __ldg(fr.p.subspace_info.Q + light_id);
    return 1;
}
__device__ __forceinline__ float connectRate_SOL(const DevFrame& fr, int e, int l, float lum) { return gamma_ss(fr, e, l) * lum * (float)fr.connections; }       // cuProg.h:70-73
__device__ __forceinline__ float3 connectRate_SOL3(const DevFrame& fr, int e, int l, float3 lum) { return gamma_ss(fr, e, l) * lum * (float)fr.connections; }   // cuProg.h:75-78
__device__ __forceinline__ float getRR(const Vtx& v) { return fmaxf(fmax3(v.color), 0.3f); }   // rmis.h:28-40 (MIN_RR_RATE .3)

__device__ __forceinline__ float getLast_pdf(const DevFrame& fr, const Vtx& Mid, float3 in_dir) {   // rmis.h:41-51
    const Pbr mat = vertex_mat(fr, Mid);
    const float3 out_vec = Mid.lastPosition - Mid.position;
    const float3 out_dir = normalize(out_vec);
    float pdf = Mid.isLastVertex_direction ? bsdf_pdf(mat, Mid.normal, in_dir, out_dir)
                                           : bsdf_pdf(mat, Mid.normal, in_dir, out_dir) / dot(out_vec, out_vec) * Mid.lastNormalProjection;
    pdf *= getRR(Mid);
    return pdf;
}
__device__ __forceinline__ float getLL_pdf(const DevFrame& fr, const Vtx& Mid, const Vtx& Last) {   // rmis.h:52-57
    const float3 in_dir = normalize(Mid.position - Last.position);
    return getLast_pdf(fr, Last, in_dir);
}
// Cross labels.  The recursive MIS classifies a light vertex with the EYE tree (tracing_weight_light) and an eye vertex with the
// LIGHT tree (tracing_weight_eye).  Both depend on the vertex alone, yet the reference walks the tree at every use: once per
// connection and again when the path is extended.  The wavefront passes compute each label once per vertex (k_eye_sample,
// k_lvc_xlabel) and hand it in through `xlabel`; a negative value means "not cached, walk the tree" -- the same function on the
// same arguments, so the result is identical either way.
__device__ __forceinline__ float tracing_weight_light(const DevFrame& fr, const Vtx& Mid, const Vtx& Last, int xlabel = -1) {   // rmis.h:58-79
    if (Last.lastBrdf || Last.isBrdf) return 0.0f;
    const int eye_label = xlabel >= 0 ? xlabel : eye_tree_label(fr, Last.position, Last.normal);
    const int light_label = Last.lastZoneId;
    const float lum_sum = Last.last_lum;
    return connectRate_SOL(fr, eye_label, light_label, lum_sum);
}
__device__ __forceinline__ void tracing_update_light(const DevFrame& fr, Vtx& Mid, const Vtx& Last) {   // rmis.h:80-95
    const float LL_pdf = getLL_pdf(fr, Mid, Last);
    const float weight = tracing_weight_light(fr, Mid, Last);
    const float last_single_pdf = Last.singlePdf;
    Mid.RMIS_pointer = ((Last.RMIS_pointer * LL_pdf) + weight) / last_single_pdf;
}
__device__ __forceinline__ float3 getFluxMultiplier(const DevFrame& fr, const Vtx& v, float3 in_dir, float3 out_dir) {   // rmis.h:102-112
    const Pbr mat = vertex_mat(fr, v);
    const float3 flux_ratio = bsdf_eval(mat, v.normal, in_dir, out_dir) / (mat.brdf ? fabsf(dot(v.normal, out_dir)) : 1.0f);
    const float pdf_ratio = bsdf_pdf(mat, v.normal, in_dir, out_dir);
    const float rr = getRR(v);
    const float cos_theta = fabsf(dot(v.normal, out_dir));
    return flux_ratio * cos_theta / pdf_ratio / rr;
}
__device__ __forceinline__ float3 getFluxMultiplier(const DevFrame& fr, const Vtx& v, float3 in_dir) {   // rmis.h:113-118
    const float3 out_vec = v.lastPosition - v.position;
    return getFluxMultiplier(fr, v, in_dir, normalize(out_vec));
}
__device__ __forceinline__ float3 tracing_weight_eye(const DevFrame& fr, const Vtx& Last, int xlabel = -1) {   // rmis.h:131-151
    if (Last.lastBrdf || Last.isBrdf) return f3(0.0f);
    if (Last.depth == 1) return f3(0.0f);   // t=1 strategy disabled (readme.md:27)
    const int eye_label = Last.lastZoneId;
    const int light_label = xlabel >= 0 ? xlabel : light_tree_label(fr, Last.position, Last.normal);
    return connectRate_SOL3(fr, eye_label, light_label, f3(1.0f));
}
__device__ __forceinline__ float getPdf(const DevFrame& fr, const Vtx& begin, const Vtx& end, float3 in_dir) {   // rmis.h:153-172
    const Pbr mat = vertex_mat(fr, begin);
    const float3 out_vec = end.position - begin.position;
    const float3 out_dir = normalize(out_vec);
    float pdf = bsdf_pdf(mat, begin.normal, in_dir, out_dir) / dot(out_vec, out_vec) * fabsf(dot(out_dir, end.normal));
    pdf *= getRR(begin);
    return pdf;
}
__device__ __forceinline__ float getPdf_from_light_source(const Vtx& light, const Vtx& end) {   // rmis.h:173-188
    const float3 conn_vec = end.position - light.position;
    const float3 conn_dir = normalize(conn_vec);
#ifdef SPC_FAST_MATH
    const float pdf_angle = fabsf(dot(light.normal, conn_dir)) * (1.0f / SPC_PI_F);
#else
    const float pdf_angle = (float)((double)fabsf(dot(light.normal, conn_dir)) / SPC_PI_D);
#endif
    const float angle2a = fabsf(dot(end.normal, conn_dir)) / (dot(conn_vec, conn_vec));
    return pdf_angle * angle2a;
}
__device__ __forceinline__ void tracing_update_eye(const DevFrame& fr, Vtx& Mid, const Vtx& Last, int last_xlabel = -1) {   // rmis.h:189-203
    const float LL_pdf = getLL_pdf(fr, Mid, Last);
    const float3 weight = tracing_weight_eye(fr, Last, last_xlabel);
    const float last_single_pdf = Last.singlePdf;
    const float3 flux_multiplier = getFluxMultiplier(fr, Last, normalize(Mid.position - Last.position));
    Mid.RMIS_pointer_3 = ((Last.RMIS_pointer_3 * LL_pdf * flux_multiplier) + weight) / last_single_pdf;
}
__device__ __forceinline__ float general_connection(const DevFrame& fr, const Vtx& eye, const Vtx& light, int eye_xlabel = -1, int light_xlabel = -1) {   // rmis.h:212-247
    if (eye.isBrdf || light.isBrdf) return 0.0f;
    const float3 connect_vec = eye.position - light.position;
    const float3 connect_dir = normalize(connect_vec);
    const float3 flux = light.flux / light.pdf;
    const float LL_pdf_A = getLL_pdf(fr, light, eye);
    const float3 flux_multiplier_0 = getFluxMultiplier(fr, eye, -connect_dir);
    const float3 weight_A = tracing_weight_eye(fr, eye, eye_xlabel);
    const float3 D_A_0 = ((eye.RMIS_pointer_3 * LL_pdf_A * flux_multiplier_0) + weight_A);
    const float3 LA = normalize(light.lastPosition - light.position);
    const float pdf_A = getPdf(fr, light, eye, LA);
    const float3 flux_multiplier_1 = getFluxMultiplier(fr, light, LA, connect_dir);
    const float D_A = sum3(D_A_0 * pdf_A * flux_multiplier_1 * flux / eye.singlePdf);
    const float weight = sum3(connectRate_SOL3(fr, eye.subspaceId, light.subspaceId, flux));
    const float LL_pdf_B = getLL_pdf(fr, eye, light);
    const float weight_B = tracing_weight_light(fr, eye, light, light_xlabel);
    const float D_B_0 = (light.RMIS_pointer * LL_pdf_B) + weight_B;
    const float3 LB = normalize(eye.lastPosition - eye.position);
    const float pdf_B = getPdf(fr, eye, light, LB);
    const float D_B = D_B_0 * pdf_B / light.singlePdf;
    return weight / (weight + D_A + D_B);
}
__device__ __forceinline__ float connection_lightSource(const DevFrame& fr, const Vtx& eye, const Vtx& light, int eye_xlabel = -1) {   // rmis.h:281-313
    if (eye.isBrdf || light.isBrdf) return 0.0f;
    const float3 connect_vec = eye.position - light.position;
    const float3 connect_dir = normalize(connect_vec);
    const float3 flux = light.flux / light.pdf;
    const float LL_pdf_A = getLL_pdf(fr, light, eye);
    const float3 flux_multiplier_0 = getFluxMultiplier(fr, eye, -connect_dir);
    const float3 weight_A = tracing_weight_eye(fr, eye, eye_xlabel);
    const float3 D_A_0 = ((eye.RMIS_pointer_3 * LL_pdf_A * flux_multiplier_0) + weight_A);
    const float pdf_A = getPdf_from_light_source(light, eye);
    const float flux_multiplier_1 = SPC_PI_F;
    const float D_A = sum3(D_A_0 * pdf_A * flux_multiplier_1 * flux / eye.singlePdf);
    const float weight = sum3(connectRate_SOL3(fr, eye.subspaceId, light.subspaceId, flux));
    const float D_B_0 = light.RMIS_pointer;
    const float3 LB = normalize(eye.lastPosition - eye.position);
    const float pdf_B = getPdf(fr, eye, light, LB);
    const float D_B = D_B_0 * pdf_B / light.singlePdf;
    return weight / (weight + D_A + D_B);
}
__device__ __forceinline__ float light_hit(const DevFrame& fr, const Vtx& eye, const Vtx& light, int eye_xlabel = -1) {   // rmis.h:359-389
    const float3 connect_vec = eye.position - light.position;
    const float3 connect_dir = normalize(connect_vec);
    const float3 flux = light.flux / light.pdf;
    const float LL_pdf_A = getLL_pdf(fr, light, eye);
    const float3 flux_multiplier_0 = getFluxMultiplier(fr, eye, -connect_dir);
    const float3 weight_A = tracing_weight_eye(fr, eye, eye_xlabel);
    const float3 D_A_0 = ((eye.RMIS_pointer_3 * LL_pdf_A * flux_multiplier_0) + weight_A);
    const float pdf_A = getPdf_from_light_source(light, eye);
    const float flux_multiplier_1 = SPC_PI_F;
    const float D_A = sum3(D_A_0 * pdf_A * flux_multiplier_1 * flux / eye.singlePdf);
    float weight = sum3(connectRate_SOL3(fr, eye.subspaceId, light.subspaceId, flux));
    if (eye.isBrdf || light.isBrdf) weight = 0.0f;
    const float D_B = light.RMIS_pointer;
    const float3 LB = normalize(eye.lastPosition - eye.position);
    const float pdf_B = getPdf(fr, eye, light, LB);
    return D_B / ((weight + D_A) / pdf_B * light.singlePdf + D_B);
}

__device__ __forceinline__ bool invalid3(float3 a) {   // ISINVALIDVALUE, raygen.cu:43
    return a.x > 100000.0f || isnan(a.x) || a.y > 100000.0f || isnan(a.y) || a.z > 100000.0f || isnan(a.z);
}

// connectVertex_SPCBPT (raygen.cu:253-303): contribution * MIS weight of joining eye vertex a to light vertex b
__device__ __forceinline__ float3 connect_vertices(const DevFrame& fr, const Vtx& a, const Vtx& b, float* w_out, int a_xlabel = -1, int b_xlabel = -1) {
    const float3 connectVec = a.position - b.position;
    const float3 connectDir = normalize(connectVec);
    const float G = fabsf(dot(a.normal, connectDir)) * fabsf(dot(b.normal, connectDir)) / dot(connectVec, connectVec);
    const float3 LA_DIR = normalize(a.lastPosition - a.position);
    const float3 LB_DIR = normalize(b.lastPosition - b.position);
    float3 fa, fb;
    const Pbr mat_a = vertex_mat(fr, a);
    fa = bsdf_eval(mat_a, a.normal, -connectDir, LA_DIR) / (mat_a.brdf ? fabsf(dot(a.normal, connectDir)) : 1.0f);
    if (!b.isOrigin) {
        const Pbr mat_b = vertex_mat(fr, b);
        fb = bsdf_eval(mat_b, b.normal, connectDir, LB_DIR) / (mat_b.brdf ? fabsf(dot(b.normal, connectDir)) : 1.0f);
    } else {
        if (dot(b.normal, -connectDir) > 0.0f) fb = f3(0.0f);
        else fb = f3(1.0f);
    }
    const float3 contri = a.flux * b.flux * fa * fb * G;
    const float pdf = a.pdf * b.pdf;
    const float w = (b.depth == 0 ? connection_lightSource(fr, a, b, a_xlabel) : general_connection(fr, a, b, a_xlabel, b_xlabel));
    if (w_out) *w_out = w;
    const float3 ans = contri / pdf * w;
    return invalid3(ans) ? f3(0.0f) : ans;
}

// A connection whose endpoints face away from each other contributes exactly zero whatever the shadow ray finds: connect_vertices
// multiplies by fa = Eval(a, N_a, -connectDir, .) and fb = Eval(b, N_b, connectDir, .) (or the one-sided emitter term), and Eval
// returns 0 when dot(N, L) <= 0 (cuProg.h:741-742).  The reference traces the shadow ray first and finds out afterwards
// (raygen.cu:405-413); here the same two dot products -- the very expressions connect_vertices / bsdf_eval evaluate -- are checked
// before the ray is queued.  A culled connection adds +0 to the pixel, as it would have: frames are bit-identical.
__device__ __forceinline__ bool connection_is_dead(float3 a_position, float3 a_normal, float3 b_position, float3 b_normal, bool b_is_origin) {
    const float3 connectVec = a_position - b_position;
    const float3 connectDir = normalize(connectVec);
    if (dot(a_normal, -connectDir) <= 0.0f) return true;                 // fa == 0
    if (b_is_origin) return dot(b_normal, -connectDir) > 0.0f;            // fb == 0 (one-sided emitter, raygen.cu:280-290)
    return dot(b_normal, connectDir) <= 0.0f;                             // fb == 0
}

// binary_sample (cuProg.h:245-264): the reference's own bisect; returns l and its pmf
__device__ __forceinline__ int binary_sample(const float* __restrict__ cmf, int size, uint32_t& seed, float& pmf) {
    const float index = rnd(seed) * 1.0f;
    int mid = size / 2 - 1, l = 0, r = size;
    while (r - l > 1) {
        if (index < __ldg(cmf + mid)) r = mid + 1;
        else l = mid + 1;
        mid = (l + r) / 2 - 1;
    }
    pmf = l == 0 ? __ldg(cmf + l) : __ldg(cmf + l) - __ldg(cmf + l - 1);
    return l;
}

// Guide tables (the "cutpoint method" for inverting a discrete CDF).  binary_sample returns, for a non-decreasing table, the first
// index l with u < cmf[l] (n-1 when there is none): ~log2(n) dependent probes.  A guide table G over the same n entries,
//     G[j] = first i with cmf[i] > j/n  (n when there is none),   j = 0..n,
// brackets that index: with j/n <= u < (j+1)/n the answer lies in [G[j], G[j+1]], usually one or two entries wide, so a search costs two
// adjacent table reads plus one or two probes.  The comparisons against cmf are the very same `u < cmf[i]`, hence the same index for
// every u; tables are only built for arrays this library produced itself (monotone by construction: running fp32 sums of non-negative
// weights divided by their total; CDF rows of positive entries), anything else keeps the reference's bisect.
__device__ __forceinline__ float guide_cut(int j, float fn) { return __fdiv_rn((float)j, fn); }
// the cell j with cut(j) <= u < cut(j+1), exact in fp32 (the product u*n may round across a cut)
__device__ __forceinline__ int guide_cell(float u, int n) {
    const float fn = (float)n;
    int j = min((int)(u * fn), n - 1);
    while (j > 0 && u < guide_cut(j, fn)) j--;
    while (j < n - 1 && !(u < guide_cut(j + 1, fn))) j++;
    return j;
}
// binary_sample through a guide table: one draw, same index and pmf
__device__ __forceinline__ int guided_sample(const float* __restrict__ cmf, const int* __restrict__ G, int size, uint32_t& seed, float& pmf) {
    const float u = rnd(seed) * 1.0f;
    const int cell = guide_cell(u, size);
    int hi = min(__ldg(G + cell + 1), size - 1);
    int lo = min(__ldg(G + cell), hi);
    while (lo < hi) {   // first i in [lo, hi) with u < cmf[i], else hi
        const int mid = (lo + hi) >> 1;
        if (u < __ldg(cmf + mid)) hi = mid;
        else lo = mid + 1;
    }
    pmf = lo == 0 ? __ldg(cmf + lo) : __ldg(cmf + lo) - __ldg(cmf + lo - 1);
    return lo;
}
// builder: one table, cooperatively by the calling block
__device__ __forceinline__ void guide_build_table(const float* __restrict__ cmf, int n, int* __restrict__ G) {
    const float fn = (float)n;
    for (int j = threadIdx.x; j <= n; j += blockDim.x) {
        const float c = guide_cut(j, fn);
        int lo = 0, hi = n;   // first i in [0, n) with cmf[i] > c, else n
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cmf[mid] > c) hi = mid;
            else lo = mid + 1;
        }
        G[j] = lo;
    }
}

// ---------------------------------------------------------------------------------------------
// the two surface programs: __closesthit__eyeSubpath (hit_program.cu:246-340) and
// __closesthit__lightSubpath (:341-438).  `Last` = path.currentVertex() before the hit, `pre_flux` /
// `pre_singlePdf` = the values the previous hit pre-loaded into path.nextVertex() (:286-287, :335).
// Outputs: Mid (the new vertex), the sampled continuation (dir, pre-loads) and `done`.
// ---------------------------------------------------------------------------------------------
struct SurfaceOut {
    float3 dir;            // next ray direction (prd.ray_direction)
    float3 next_flux;      // pre-load for the next vertex
    float  next_singlePdf;
    bool   done;
};
__device__ __forceinline__ void surface_hit(const DevFrame& fr, const Vtx& Last, float3 pre_flux, float pre_singlePdf, const LocalGeom& geom,
                                            float t_hit, float3 ray_direction, bool light_side, uint32_t& seed, Vtx& Mid, SurfaceOut& out,
                                            int last_xlabel = -1, bool label_later = false) {
    const float3 inver_ray_direction = -ray_direction;
    const Pbr currentPbr = shade_pbr(fr.sc, geom.material, geom.uv);
    float3 N = geom.Ng;
    if (dot(N, ray_direction) > 0.f) N = -N;
    out.dir = bsdf_sample(currentPbr, N, inver_ray_direction, seed);
    const float pdf = bsdf_pdf(currentPbr, N, inver_ray_direction, out.dir);
    out.done = !(pdf > 0.0f);

    vtx_zero(Mid);
    Mid.position = geom.P;
    Mid.normal = N;
    Mid.type = SPC_VTYPE_NORMALHIT;
    const float pdf_G = fabsf(dot(Mid.normal, ray_direction) * dot(Last.normal, ray_direction)) / (t_hit * t_hit);
    if (Last.isOrigin) Mid.flux = Last.flux * pdf_G;
    else Mid.flux = pre_flux * Last.flux * pdf_G;
    out.next_flux = bsdf_eval(currentPbr, N, -ray_direction, out.dir) / (currentPbr.brdf ? fabsf(dot(Mid.normal, out.dir)) : 1.0f);
    out.next_singlePdf = pdf;
    Mid.lastPosition = Last.position;
    Mid.color = currentPbr.base_color;
    Mid.lastNormalProjection = fabsf(dot(Last.normal, ray_direction));
    Mid.materialId = (short)geom.material;
    // label_later (wavefront eye pass): k_eye_sample classifies the vertex (both trees in one lockstep walk) and patches the field
    Mid.subspaceId = label_later ? (short)-1
                                 : (short)(light_side ? light_tree_label(fr, Mid.position, Mid.normal) : eye_tree_label(fr, Mid.position, Mid.normal));
    Mid.lastZoneId = Last.subspaceId;
    Mid.lastBrdf = Last.isBrdf;
    Mid.isOrigin = 0;
    Mid.depth = (short)(Last.depth + 1);
    Mid.uv = geom.uv;
    Mid.singlePdf = pre_singlePdf * pdf_G / fabsf(dot(Last.normal, ray_direction));
    Mid.pdf = Last.pdf * Mid.singlePdf;
    if (light_side) Mid.last_lum = sum3(Last.flux / Last.pdf);
    Mid.lastSinglePdf = Last.singlePdf;
    Mid.isLastVertex_direction = 0;
    if (light_side) {
        if (Last.isOrigin) Mid.RMIS_pointer = Last.RMIS_pointer / Last.singlePdf;   // rmis::tracing_init_light, rmis.h:22-26
        else tracing_update_light(fr, Mid, Last);
    } else {
        if (Mid.depth == 1) Mid.RMIS_pointer_3 = f3(0.0f);                          // rmis::tracing_init_eye, rmis.h:204-207
        else tracing_update_eye(fr, Mid, Last, last_xlabel);
    }
    const float r = rnd(seed);
    float rr_rate = fmax3(Mid.color);
    rr_rate = rr_rate < 0.3f ? 0.3f : rr_rate;   // RR_MIN_LIMIT / MIN_RR_RATE (optixPathTracer.h:34-35)
    if (r > rr_rate) out.done = true;
    else out.next_singlePdf *= rr_rate;
}

// __closesthit__eyeSubpath_LightSource (hit_program.cu:62-147).  Returns false when the emitter is seen
// from behind (no vertex is added).
__device__ __forceinline__ bool eye_hits_light(const DevFrame& fr, const Vtx& Last, float3 pre_flux, float pre_singlePdf, const LocalGeom& geom,
                                               float t_hit, float3 ray_direction, Vtx& Mid, int last_xlabel = -1) {
    const spc_light& light = fr.sc.lights[geom.light];
    const float3 ln = ld3(light.normal);
    if (dot(ray_direction, ln) > 0) return false;
    vtx_zero(Mid);
    Mid.position = geom.P;
    Mid.normal = ln;
    Mid.type = SPC_VTYPE_HIT_LIGHT_SOURCE;
    Mid.uv = geom.uv;
    LightSample ls;
    light_reverse_sample(fr, geom.light, Mid.uv.x, Mid.uv.y, ls);
    const float lightPdf = ls.pdf;
    const float pdf_G = fabsf(dot(Mid.normal, ray_direction) * dot(Last.normal, ray_direction)) / (t_hit * t_hit);
    if (Last.isOrigin) Mid.flux = Last.flux * pdf_G * ls.emission;
    else Mid.flux = pre_flux * Last.flux * pdf_G * ls.emission;
    Mid.lastPosition = Last.position;
    Mid.lastNormalProjection = fabsf(dot(Last.normal, ray_direction));
    Mid.subspaceId = (short)ls.subspaceId;
    Mid.lastZoneId = Last.subspaceId;
    Mid.singlePdf = pre_singlePdf * pdf_G / fabsf(dot(Last.normal, ray_direction));
    Mid.pdf = Last.pdf * Mid.singlePdf;
    Mid.materialId = (short)geom.light;
    Mid.depth = (short)(Last.depth + 1);
    if (Mid.depth == 1) {
        Mid.RMIS_pointer = 1.0f;
        return true;
    }
    Vtx virtual_light;
    vtx_zero(virtual_light);
    virtual_light.type = SPC_VTYPE_QUAD;
    virtual_light.position = Mid.position;
    virtual_light.RMIS_pointer = 1;
    virtual_light.normal = Mid.normal;
    virtual_light.pdf = lightPdf;
    virtual_light.singlePdf = lightPdf;
    virtual_light.flux = ls.emission;
    virtual_light.subspaceId = Mid.subspaceId;
    virtual_light.isBrdf = 0;
    Mid.RMIS_pointer = cm_div64(1.0f, light_hit(fr, Last, virtual_light, last_xlabel));
    return true;
}

}  // namespace spc
