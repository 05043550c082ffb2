// orc_render.cpp -- CPU oracle of the SPCBPT render path (see orc_render.h).  TEST INFRASTRUCTURE ONLY.
#include "orc_render.h"
#include "orc_internal.h"
#include <algorithm>
#include <cfloat>
#include <cstring>

namespace orc {

int g_jitter_rtl = 0;
static const float PIf = 3.14159265358979323846f;   // M_PIf, sutil/vec_math.h
static const double PId = 3.14159265358979323846;   // M_PI
static const float SCENE_EPS = 1e-3f;               // cuProg.h:39

static inline float absf(float x) { return std::fabs(x); }
static inline f3 operator+(f3 a, float b) { return f3{a.x + b, a.y + b, a.z + b}; }   // sutil/vec_math.h float3+float

// ============================================================================================
// materials + textures
// ============================================================================================
// fp32 bilinear fetch, wrap addressing, unnormalised texel centres at +0.5: the addressing of the
// reference's samplers (cudaAddressModeWrap + cudaFilterModeLinear, scene_shift.cpp:57-60), with
// fp32 weights instead of CUDA's 9-bit ones.  Same code as ref_shim_tex2D and csrc/shade.cuh tex_fetch.
static void tex_fetch(const Texture& t, float u, float v, float out[4]) {
    const float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    const float fx = std::floor(x), fy = std::floor(y);
    const float ax = x - fx, ay = y - fy;
    auto wrap = [](int i, int n) { i %= n; return i < 0 ? i + n : i; };
    const int x0 = wrap((int)fx, t.w), x1 = wrap((int)fx + 1, t.w), y0 = wrap((int)fy, t.h), y1 = wrap((int)fy + 1, t.h);
    for (int c = 0; c < 4; c++) {
        const float t00 = t.rgba[4 * ((size_t)y0 * t.w + x0) + c] * (1.0f / 255.0f);
        const float t10 = t.rgba[4 * ((size_t)y0 * t.w + x1) + c] * (1.0f / 255.0f);
        const float t01 = t.rgba[4 * ((size_t)y1 * t.w + x0) + c] * (1.0f / 255.0f);
        const float t11 = t.rgba[4 * ((size_t)y1 * t.w + x1) + c] * (1.0f / 255.0f);
        const float a = t00 + ax * (t10 - t00), b = t01 + ax * (t11 - t01);
        out[c] = a + ay * (b - a);
    }
}

Pbr load_pbr(const Scene& sc, int id) {
    const spc_pbr& s = sc.materials[id];
    Pbr m;
    m.base_color = f3{s.base_color[0], s.base_color[1], s.base_color[2]};
    m.metallic = s.metallic; m.roughness = s.roughness; m.specular = s.specular; m.specularTint = s.specularTint;
    m.subsurface = s.subsurface; m.sheen = s.sheen; m.sheenTint = s.sheenTint; m.clearcoat = s.clearcoat;
    m.clearcoatGloss = s.clearcoatGloss; m.brdf = s.brdf != 0;
    return m;
}

// ColorTexSample (hit_program.cu:182-198) + sampleTexture (src/cuda/LocalShading.h:37-53) +
// linearize (cuProg.h:361-368).  RoughnessAndMetallicTexSample (hit_program.cu:199-209) multiplies by
// 1 because no scene sets metallic_roughness_tex (scene_shift.cpp:64-91).
static Pbr shade_pbr(const Scene& sc, int id, float uvx, float uvy) {
    Pbr m = load_pbr(sc, id);
    const spc_texture_ref& tr = sc.materials[id].base_color_tex;
    if (tr.tex != 0) {
        const float sx = uvx * tr.texcoord_scale[0], sy = uvy * tr.texcoord_scale[1];
        const float rx = tr.texcoord_rotation[0], ry = tr.texcoord_rotation[1];
        const float tu = (sx * ry + sy * rx) + tr.texcoord_offset[0];
        const float tv = (sx * (-rx) + sy * ry) + tr.texcoord_offset[1];
        float c[4];
        tex_fetch(sc.textures[(size_t)tr.tex - 1], tu, tv, c);
        m.base_color = f3{cm_powf(c[0], 2.2f), cm_powf(c[1], 2.2f), cm_powf(c[2], 2.2f)};
    }
    m.roughness *= 1.0f;
    m.metallic *= 1.0f;
    return m;
}

// ============================================================================================
// Disney BSDF (cuProg.h:684-899)
// ============================================================================================
static inline float sqr(float x) { return x * x; }
static inline float SchlickFresnel(float u) {              // cuProg.h:686-691
    const float m = clampf(1.0f - u, 0.0f, 1.0f);
    const float m2 = m * m;
    return m2 * m2 * m;
}
static inline float GTR1(float NDotH, float a) {            // cuProg.h:693-699
    if (a >= 1.0f) return (1.0f / PIf);
    const float a2 = a * a;
    const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return (a2 - 1.0f) / (PIf * cm_logf(a2) * t);
}
static inline float GTR2(float NDotH, float a) {            // cuProg.h:701-706
    const float a2 = a * a;
    const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return a2 / (PIf * t * t);
}
static inline float smithG_GGX(float NDotv, float alphaG) { // cuProg.h:708-713
    const float a = alphaG * alphaG;
    const float b = NDotv * NDotv;
    return 1.0f / (NDotv + std::sqrt(a + b - a * b));
}

f3 bsdf_eval(const Pbr& mat, f3 N, f3 V, f3 L) {            // Tracer::Eval, cuProg.h:735-799
    const float NDotL = dot(N, L);
    const float NDotV = dot(N, V);
    if (NDotL <= 0.0f || NDotV <= 0.0f) return mk3(0.0f);
    const f3 H = normalize(L + V);
    const float NDotH = dot(N, H);
    const float LDotH = dot(L, H);
    const f3 Cdlin = mat.base_color;
    const float Cdlum = 0.3f * Cdlin.x + 0.6f * Cdlin.y + 0.1f * Cdlin.z;
    const f3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : mk3(1.0f);
    const f3 Cspec0 = lerp3(mat.specular * 0.08f * lerp3(mk3(1.0f), Ctint, mat.specularTint), Cdlin, mat.metallic);
    const f3 Csheen = lerp3(mk3(1.0f), Ctint, mat.sheenTint);
    const float FL = SchlickFresnel(NDotL), FV = SchlickFresnel(NDotV);
    const float Fd90 = 0.5f + 2.0f * LDotH * LDotH * mat.roughness;
    const float Fd = lerpf(1.0f, Fd90, FL) * lerpf(1.0f, Fd90, FV);
    const float Fss90 = LDotH * LDotH * mat.roughness;
    const float Fss = lerpf(1.0f, Fss90, FL) * lerpf(1.0f, Fss90, FV);
    const float ss = 1.25f * (Fss * (1.0f / (NDotL + NDotV) - 0.5f) + 0.5f);
    const float a = std::fmax(0.001f, mat.roughness);
    const float Ds = GTR2(NDotH, a);
    const float FH = SchlickFresnel(LDotH);
    const f3 Fs = lerp3(Cspec0, mk3(1.0f), FH);
    const float roughg = sqr(mat.roughness * 0.5f + 0.5f);
    const float Gs = smithG_GGX(NDotL, roughg) * smithG_GGX(NDotV, roughg);
    const f3 Fsheen = FH * mat.sheen * Csheen;
    const float Dr = GTR1(NDotH, lerpf(0.1f, 0.001f, mat.clearcoatGloss));
    const float Fr = lerpf(0.04f, 1.0f, FH);
    const float Gr = smithG_GGX(NDotL, 0.25f) * smithG_GGX(NDotV, 0.25f);
    const f3 out = ((1.0f / PIf) * lerpf(Fd, ss, mat.subsurface) * Cdlin + Fsheen) * (1.0f - mat.metallic) + Gs * Fs * Ds +
                   0.25f * mat.clearcoat * Gr * Fr * Dr;
    return out;
}

struct Onb {                                                  // cuProg.h:81-112
    f3 t, b, n;
    explicit Onb(f3 normal) {
        n = normal;
        if (absf(n.x) > absf(n.z)) b = f3{-n.y, n.x, 0.f};
        else b = f3{0.f, -n.z, n.y};
        b = normalize(b);
        t = cross(b, n);
    }
    f3 inverse_transform(f3 p) const { return p.x * t + p.y * b + p.z * n; }
};
static inline f3 cosine_sample_hemisphere(float u1, float u2) {   // cuProg.h:114-124
    const float r = std::sqrt(u1);
    const float phi = 2.0f * PIf * u2;
    f3 p;
    p.x = r * cm_cosf(phi);
    p.y = r * cm_sinf(phi);
    p.z = std::sqrt(std::fmax(0.0f, 1.0f - p.x * p.x - p.y * p.y));
    return p;
}

f3 bsdf_sample(const Pbr& mat, f3 N, f3 V, uint32_t& seed) {   // Tracer::Sample, cuProg.h:826-866
    f3 dir;
    const float probability = rnd(seed);
    const float diffuseRatio = 0.5f * (1.0f - mat.metallic);
    const float r1 = rnd(seed);
    const float r2 = rnd(seed);
    const Onb onb(N);
    if (probability < diffuseRatio) {
        dir = onb.inverse_transform(cosine_sample_hemisphere(r1, r2));
    } else {
        const float a = std::fmax(0.001f, mat.roughness);
        const float phi = r1 * 2.0f * PIf;
        const float cosTheta = std::sqrt((1.0f - r2) / (1.0f + (a * a - 1.0f) * r2));
        const float sinTheta = std::sqrt(1.0f - (cosTheta * cosTheta));
        const float sinPhi = cm_sinf(phi);
        const float cosPhi = cm_cosf(phi);
        f3 half = f3{sinTheta * cosPhi, sinTheta * sinPhi, cosTheta};
        half = onb.inverse_transform(half);
        dir = 2.0f * dot(V, half) * half - V;
    }
    return dir;
}

float bsdf_pdf(const Pbr& mat, f3 n, f3 V, f3 L) {              // Tracer::Pdf, cuProg.h:868-899
    const float specularAlpha = std::fmax(0.001f, mat.roughness);
    const float clearcoatAlpha = lerpf(0.1f, 0.001f, mat.clearcoatGloss);
    const float diffuseRatio = 0.5f * (1.f - mat.metallic);
    const float specularRatio = 1.f - diffuseRatio;
    const f3 half = normalize(L + V);
    const float cosTheta = absf(dot(half, n));
    const float pdfGTR2 = GTR2(cosTheta, specularAlpha) * cosTheta;
    const float pdfGTR1 = GTR1(cosTheta, clearcoatAlpha) * cosTheta;
    const float ratio = 1.0f / (1.0f + mat.clearcoat);
    // `/ (4.0 * abs(...))`: the literal is a double, so this one division is done in fp64 (cuProg.h:892)
    const float pdfSpec = (float)((double)lerpf(pdfGTR1, pdfGTR2, ratio) / (4.0 * (double)absf(dot(L, half))));
    const float pdfDiff = absf(dot(L, n)) * (1.0f / PIf);
    return diffuseRatio * pdfDiff + specularRatio * pdfSpec;
}

// ============================================================================================
// classification: classTree::tree_index (decisionTree/classTree_common.h:39-51), reached through
// labelUnit::getLabel (cuProg.h:1109-1123): null tree -> label 0; DIR_JUDGE 0 -> direction unused.
// ============================================================================================
int tree_label(const spc_tree_node* root, f3 position, f3 normal) {
    if (!root) return 0;
    int node = 0;
    while (!root[node].leaf) {
        const spc_tree_node& nd = root[node];
        const f3 q = nd.type == 0 ? position : (nd.type == 1 ? normal : mk3(0.0f));
        int ind = 0;
        ind += q.x > nd.mid.x ? 1 : 0;
        ind += q.y > nd.mid.y ? 2 : 0;
        ind += q.z > nd.mid.z ? 4 : 0;
        node = nd.child[ind];
    }
    return root[node].label;
}

// ============================================================================================
// hit-point reconstruction: getLocalGeometry (src/cuda/LocalGeometry.h:59-176), identity transforms
// ============================================================================================
struct LocalGeom { f3 P, Ng; float uvx, uvy; };
static LocalGeom local_geometry(const Tri& tr, float bu, float bv) {
    LocalGeom g;
    const float w = 1.0f - bu - bv;
    g.P = w * tr.v0 + bu * tr.v1 + bv * tr.v2;
    g.Ng = normalize(cross(tr.v1 - tr.v0, tr.v2 - tr.v0));
    g.uvx = w * tr.uv[0][0] + bu * tr.uv[1][0] + bv * tr.uv[2][0];
    g.uvy = w * tr.uv[0][1] + bu * tr.uv[1][1] + bv * tr.uv[2][1];
    return g;
}

// ============================================================================================
// light sampling: Tracer::lightSample (cuProg.h:554-666), QUAD lights only
// ============================================================================================
void light_reverse_sample(const Frame& fr, const spc_light& L, float r1, float r2, LightSample& s) {   // cuProg.h:571-601
    s.light = &L;
    const float r3 = 1 - r1 - r2;
    s.position = ld(L.u) * r1 + ld(L.v) * r2 + ld(L.corner) * r3;
    s.emission = ld(L.emission);
    s.pdf = (float)(1.0 / (double)L.area);
    s.pdf /= (float)(unsigned)fr.sc->lights.size();
    s.uvx = r1; s.uvy = r2;
    const int xb = std::max(0, std::min((int)std::floor(s.uvx * L.divLevel), L.divLevel - 1));
    const int yb = std::max(0, std::min((int)std::floor(s.uvy * L.divLevel), L.divLevel - 1));
    const int lightSpaceId = L.ssBase + xb * L.divLevel + yb;
    s.subspaceId = fr.K - lightSpaceId - 1;
}
void light_sample_pos(const Frame& fr, const spc_light& L, uint32_t& seed, LightSample& s) {           // cuProg.h:602-621
    const float r1 = rnd(seed);
    const float r2 = rnd(seed);
    light_reverse_sample(fr, L, r1, r2, s);
}
static void light_trace_mode(LightSample& s, uint32_t& seed) {                                                // cuProg.h:648-665
    const float r1 = rnd(seed);
    const float r2 = rnd(seed);
    const Onb onb(ld(s.light->normal));
    s.direction = onb.inverse_transform(cosine_sample_hemisphere(r1, r2));
    s.dir_pdf = absf(dot(s.direction, ld(s.light->normal))) / PIf;
}
int pick_light(const Frame& fr, uint32_t& seed) {                                                      // raygen.cu:639, cuProg.h:624
    const int n = (int)fr.sc->lights.size();
    return std::max(0, std::min((int)std::floor(rnd(seed) * (float)(unsigned)n), n - 1));
}
void init_vertex_from_light_sample(const LightSample& s, spc_vertex& v) {                              // raygen.cu:172-195
    st(v.position, s.position);
    st(v.normal, ld(s.light->normal));
    st(v.flux, s.emission);
    v.pdf = s.pdf;
    v.singlePdf = v.pdf;
    v.isOrigin = 1;
    v.isBrdf = 0;
    v.subspaceId = (int16_t)s.subspaceId;
    v.depth = 0;
    v.materialId = (int16_t)s.light->id;
    v.RMIS_pointer = 1;
    v.uv.x = s.uvx; v.uv.y = s.uvy;
    v.type = SPC_VTYPE_QUAD;
}

// ============================================================================================
// recursive MIS (rmis.h)
// ============================================================================================
Pbr vertex_mat(const Frame& fr, const spc_vertex& v) {                       // rmis::getMat, rmis.h:16-21
    Pbr m = load_pbr(*fr.sc, v.materialId);
    m.base_color = ld(v.color);
    return m;
}
static inline float Gamma(const Frame& fr, int eye_id, int light_id) {               // optixPathTracer.h:173-181
    const float* C = fr.p.subspace_info.CMFGamma;
    if (C && fr.p.subspace_info.Q)
        return light_id == 0 ? C[eye_id * fr.K + light_id] : C[eye_id * fr.K + light_id] - C[eye_id * fr.K + light_id - 1];
    return 1;
}
static inline float gamma_ss(const Frame& fr, int eye_id, int light_id) {            // optixPathTracer.h:182-189
    if (fr.p.subspace_info.CMFGamma && fr.p.subspace_info.Q) return Gamma(fr, eye_id, light_id) / fr.p.subspace_info.Q[light_id];
    return 1;
}
static inline float connectRate_SOL(const Frame& fr, int e, int l, float lum) { return gamma_ss(fr, e, l) * lum * (float)fr.connections; }   // cuProg.h:70-73
static inline f3 connectRate_SOL3(const Frame& fr, int e, int l, f3 lum) { return gamma_ss(fr, e, l) * lum * (float)fr.connections; }       // cuProg.h:75-78

static inline float getRR(const spc_vertex& v) { return std::fmax(fmax3(ld(v.color)), 0.3f); }   // rmis.h:28-40 (RR_MIN_LIMIT, MIN_RR_RATE .3)

static float getLast_pdf(const Frame& fr, const spc_vertex& Mid, f3 in_dir) {         // rmis.h:41-51
    const Pbr mat = vertex_mat(fr, Mid);
    const f3 out_vec = ld(Mid.lastPosition) - ld(Mid.position);
    const f3 out_dir = normalize(out_vec);
    float pdf = Mid.isLastVertex_direction
                    ? bsdf_pdf(mat, ld(Mid.normal), in_dir, out_dir)
                    : bsdf_pdf(mat, ld(Mid.normal), in_dir, out_dir) / dot(out_vec, out_vec) * Mid.lastNormalProjection;
    pdf *= getRR(Mid);
    return pdf;
}
static float getLL_pdf(const Frame& fr, const spc_vertex& Mid, const spc_vertex& Last) {   // rmis.h:52-57
    const f3 in_dir = normalize(ld(Mid.position) - ld(Last.position));
    return getLast_pdf(fr, Last, in_dir);
}
static float tracing_weight_light(const Frame& fr, const spc_vertex& Mid, const spc_vertex& Last) {   // rmis.h:58-79
    if (Last.lastBrdf || Last.isBrdf) return 0.0f;
    const f3 inver_dir = normalize(ld(Mid.position) - ld(Last.position));
    (void)inver_dir;
    const int eye_label = tree_label(fr.p.subspace_info.eye_tree, ld(Last.position), ld(Last.normal));
    const int light_label = Last.lastZoneId;
    const float lum_sum = Last.last_lum;
    return connectRate_SOL(fr, eye_label, light_label, lum_sum);
}
static void tracing_init_light(spc_vertex& Mid, const spc_vertex& Last) { Mid.RMIS_pointer = Last.RMIS_pointer / Last.singlePdf; }   // rmis.h:22-26
static void tracing_update_light(const Frame& fr, spc_vertex& Mid, const spc_vertex& Last) {          // rmis.h:80-95
    const float LL_pdf = getLL_pdf(fr, Mid, Last);
    const float weight = tracing_weight_light(fr, Mid, Last);
    const float last_single_pdf = Last.singlePdf;
    Mid.RMIS_pointer = ((Last.RMIS_pointer * LL_pdf) + weight) / last_single_pdf;
}
static f3 getFluxMultiplier(const Frame& fr, const spc_vertex& v, f3 in_dir, f3 out_dir) {            // rmis.h:102-112
    const Pbr mat = vertex_mat(fr, v);
    const f3 flux_ratio = bsdf_eval(mat, ld(v.normal), in_dir, out_dir) / (mat.brdf ? absf(dot(ld(v.normal), out_dir)) : 1.0f);
    const float pdf_ratio = bsdf_pdf(mat, ld(v.normal), in_dir, out_dir);
    const float rr = getRR(v);
    const float cos_theta = absf(dot(ld(v.normal), out_dir));
    return flux_ratio * cos_theta / pdf_ratio / rr;
}
static f3 getFluxMultiplier(const Frame& fr, const spc_vertex& v, f3 in_dir) {                        // rmis.h:113-118
    const f3 out_vec = ld(v.lastPosition) - ld(v.position);
    return getFluxMultiplier(fr, v, in_dir, normalize(out_vec));
}
static f3 tracing_weight_eye(const Frame& fr, const spc_vertex& Mid, const spc_vertex& Last) {        // rmis.h:131-151
    if (Last.lastBrdf || Last.isBrdf) return mk3(0.0f);
    if (Last.depth == 1) return mk3(0.0f);   // t=1 strategy disabled (readme.md:27)
    const int eye_label = Last.lastZoneId;
    const int light_label = tree_label(fr.p.subspace_info.light_tree, ld(Last.position), ld(Last.normal));
    (void)Mid;
    return connectRate_SOL3(fr, eye_label, light_label, mk3(1.0f));
}
static float getPdf(const Frame& fr, const spc_vertex& begin, const spc_vertex& end, f3 in_dir) {     // rmis.h:153-172
    const Pbr mat = vertex_mat(fr, begin);
    const f3 out_vec = ld(end.position) - ld(begin.position);
    const f3 out_dir = normalize(out_vec);
    float pdf = bsdf_pdf(mat, ld(begin.normal), in_dir, out_dir) / dot(out_vec, out_vec) * absf(dot(out_dir, ld(end.normal)));
    pdf *= getRR(begin);
    return pdf;
}
static float getPdf_from_light_source(const spc_vertex& light, const spc_vertex& end) {               // rmis.h:173-188
    const f3 conn_vec = ld(end.position) - ld(light.position);
    const f3 conn_dir = normalize(conn_vec);
    const float pdf_angle = (float)((double)absf(dot(ld(light.normal), conn_dir)) / PId);
    const float angle2a = absf(dot(ld(end.normal), conn_dir)) / (dot(conn_vec, conn_vec));
    return pdf_angle * angle2a;
}
static void tracing_update_eye(const Frame& fr, spc_vertex& Mid, const spc_vertex& Last) {            // rmis.h:189-203
    const float LL_pdf = getLL_pdf(fr, Mid, Last);
    const f3 weight = tracing_weight_eye(fr, Mid, Last);
    const float last_single_pdf = Last.singlePdf;
    const f3 flux_multiplier = getFluxMultiplier(fr, Last, normalize(ld(Mid.position) - ld(Last.position)));
    st(Mid.RMIS_pointer_3, ((ld(Last.RMIS_pointer_3) * LL_pdf * flux_multiplier) + weight) / last_single_pdf);
}
static float general_connection(const Frame& fr, const spc_vertex& eye, const spc_vertex& light) {    // rmis.h:212-247
    if (eye.isBrdf || light.isBrdf) return 0.0f;
    const f3 connect_vec = ld(eye.position) - ld(light.position);
    const f3 connect_dir = normalize(connect_vec);
    const f3 flux = ld(light.flux) / light.pdf;
    const float LL_pdf_A = getLL_pdf(fr, light, eye);
    const f3 flux_multiplier_0 = getFluxMultiplier(fr, eye, -connect_dir);
    const f3 weight_A = tracing_weight_eye(fr, light, eye);
    const f3 D_A_0 = ((ld(eye.RMIS_pointer_3) * LL_pdf_A * flux_multiplier_0) + weight_A);
    const f3 LA = normalize(ld(light.lastPosition) - ld(light.position));
    const float pdf_A = getPdf(fr, light, eye, LA);
    const f3 flux_multiplier_1 = getFluxMultiplier(fr, light, LA, connect_dir);
    const float D_A = sum3(D_A_0 * pdf_A * flux_multiplier_1 * flux / eye.singlePdf);
    const float weight = sum3(connectRate_SOL3(fr, eye.subspaceId, light.subspaceId, flux));
    const float LL_pdf_B = getLL_pdf(fr, eye, light);
    const float weight_B = tracing_weight_light(fr, eye, light);
    const float D_B_0 = (light.RMIS_pointer * LL_pdf_B) + weight_B;
    const f3 LB = normalize(ld(eye.lastPosition) - ld(eye.position));
    const float pdf_B = getPdf(fr, eye, light, LB);
    const float D_B = D_B_0 * pdf_B / light.singlePdf;
    return weight / (weight + D_A + D_B);
}
static float connection_lightSource(const Frame& fr, const spc_vertex& eye, const spc_vertex& light) {   // rmis.h:281-313
    if (eye.isBrdf || light.isBrdf) return 0.0f;
    const f3 connect_vec = ld(eye.position) - ld(light.position);
    const f3 connect_dir = normalize(connect_vec);
    const f3 flux = ld(light.flux) / light.pdf;
    const float LL_pdf_A = getLL_pdf(fr, light, eye);
    const f3 flux_multiplier_0 = getFluxMultiplier(fr, eye, -connect_dir);
    const f3 weight_A = tracing_weight_eye(fr, light, eye);
    const f3 D_A_0 = ((ld(eye.RMIS_pointer_3) * LL_pdf_A * flux_multiplier_0) + weight_A);
    const float pdf_A = getPdf_from_light_source(light, eye);
    const float flux_multiplier_1 = PIf;
    const float D_A = sum3(D_A_0 * pdf_A * flux_multiplier_1 * flux / eye.singlePdf);
    const float weight = sum3(connectRate_SOL3(fr, eye.subspaceId, light.subspaceId, flux));
    const float D_B_0 = light.RMIS_pointer;
    const f3 LB = normalize(ld(eye.lastPosition) - ld(eye.position));
    const float pdf_B = getPdf(fr, eye, light, LB);
    const float D_B = D_B_0 * pdf_B / light.singlePdf;
    return weight / (weight + D_A + D_B);
}
static float light_hit(const Frame& fr, const spc_vertex& eye, const spc_vertex& light) {               // rmis.h:359-389
    const f3 connect_vec = ld(eye.position) - ld(light.position);
    const f3 connect_dir = normalize(connect_vec);
    const f3 flux = ld(light.flux) / light.pdf;
    const float LL_pdf_A = getLL_pdf(fr, light, eye);
    const f3 flux_multiplier_0 = getFluxMultiplier(fr, eye, -connect_dir);
    const f3 weight_A = tracing_weight_eye(fr, light, eye);
    const f3 D_A_0 = ((ld(eye.RMIS_pointer_3) * LL_pdf_A * flux_multiplier_0) + weight_A);
    const float pdf_A = getPdf_from_light_source(light, eye);
    const float flux_multiplier_1 = PIf;
    const float D_A = sum3(D_A_0 * pdf_A * flux_multiplier_1 * flux / eye.singlePdf);
    float weight = sum3(connectRate_SOL3(fr, eye.subspaceId, light.subspaceId, flux));
    if (eye.isBrdf || light.isBrdf) weight = 0.0f;
    const float D_B = light.RMIS_pointer;
    const f3 LB = normalize(ld(eye.lastPosition) - ld(eye.position));
    const float pdf_B = getPdf(fr, eye, light, LB);
    return D_B / ((weight + D_A) / pdf_B * light.singlePdf + D_B);
}

// ============================================================================================
// path state: BDPTPath keeps 3 vertices in a ring (BDPTVertex.h:72-117); the programs only ever
// touch current/last/next, and `next` carries the two values pre-loaded by the previous hit
// (flux = BSDF value, singlePdf), hit_program.cu:286-287 + :335
// ============================================================================================
bool invalid3(f3 a) {                 // ISINVALIDVALUE, raygen.cu:43
    return a.x > 100000.0f || std::isnan(a.x) || a.y > 100000.0f || std::isnan(a.y) || a.z > 100000.0f || std::isnan(a.z);
}

// __closesthit__eyeSubpath (hit_program.cu:246-340) / __closesthit__lightSubpath (:341-438)
static void closesthit_surface(const Frame& fr, Payload& prd, const Tri& tr, const Hit& h, f3 ray_direction, bool light_side) {
    const LocalGeom geom = local_geometry(tr, h.u, h.v);
    const float t_hit = h.t;
    const f3 inver_ray_direction = -ray_direction;
    const Pbr currentPbr = shade_pbr(*fr.sc, tr.material, geom.uvx, geom.uvy);
    f3 N = geom.Ng;
    if (dot(N, ray_direction) > 0.f) N = -N;
    prd.ray_direction = bsdf_sample(currentPbr, N, inver_ray_direction, prd.seed);
    prd.pdf = bsdf_pdf(currentPbr, N, inver_ray_direction, prd.ray_direction);
    prd.origin = geom.P;
    if (!(prd.pdf > 0.0f)) prd.done = true;

    prd.path.size++;
    spc_vertex& Mid = prd.path.cur();
    spc_vertex& Next = prd.path.next();
    const spc_vertex& Last = prd.path.last();
    st(Mid.position, geom.P);
    st(Mid.normal, N);
    Mid.type = SPC_VTYPE_NORMALHIT;
    const float pdf_G = absf(dot(ld(Mid.normal), ray_direction) * dot(ld(Last.normal), ray_direction)) / (t_hit * t_hit);
    if (Last.isOrigin) st(Mid.flux, ld(Last.flux) * pdf_G);
    else st(Mid.flux, ld(Mid.flux) * ld(Last.flux) * pdf_G);
    st(Next.flux, bsdf_eval(currentPbr, N, -ray_direction, prd.ray_direction) / (currentPbr.brdf ? absf(dot(ld(Mid.normal), prd.ray_direction)) : 1.0f));
    Next.singlePdf = prd.pdf;
    Mid.lastPosition = Last.position;
    st(Mid.color, currentPbr.base_color);
    Mid.lastNormalProjection = absf(dot(ld(Last.normal), ray_direction));
    Mid.materialId = (int16_t)tr.material;
    Mid.subspaceId = (int16_t)tree_label(light_side ? fr.p.subspace_info.light_tree : fr.p.subspace_info.eye_tree, ld(Mid.position), ld(Mid.normal));
    Mid.lastZoneId = Last.subspaceId;
    Mid.lastBrdf = Last.isBrdf;
    Mid.isOrigin = 0;
    Mid.depth = (int16_t)(Last.depth + 1);
    Mid.uv.x = geom.uvx; Mid.uv.y = geom.uvy;
    Mid.singlePdf = Mid.singlePdf * pdf_G / absf(dot(ld(Last.normal), ray_direction));
    Mid.pdf = Last.pdf * Mid.singlePdf;
    if (light_side) Mid.last_lum = sum3(ld(Last.flux) / Last.pdf);
    Mid.lastSinglePdf = Last.singlePdf;
    Mid.isLastVertex_direction = 0;
    if (light_side) {
        if (Last.isOrigin) tracing_init_light(Mid, Last);
        else tracing_update_light(fr, Mid, Last);
    } else {
        if (Mid.depth == 1) st(Mid.RMIS_pointer_3, mk3(0.0f));   // rmis::tracing_init_eye, rmis.h:204-207
        else tracing_update_eye(fr, Mid, Last);
    }
    const float r = rnd(prd.seed);
    float rr_rate = fmax3(ld(Mid.color));
    rr_rate = rr_rate < 0.3f ? 0.3f : rr_rate;   // RR_MIN_LIMIT, MIN_RR_RATE (optixPathTracer.h:34-35)
    if (r > rr_rate) prd.done = true;
    else Next.singlePdf *= rr_rate;
}

// __closesthit__eyeSubpath_LightSource (hit_program.cu:62-147)
static void closesthit_eye_light(const Frame& fr, Payload& prd, const Tri& tr, const Hit& h, f3 ray_direction) {
    prd.done = true;
    const spc_light& light = fr.sc->lights[tr.light];
    if (dot(prd.ray_direction, ld(light.normal)) > 0) return;
    const LocalGeom geom = local_geometry(tr, h.u, h.v);
    const float t_hit = h.t;
    prd.path.size++;
    spc_vertex& Mid = prd.path.cur();
    const spc_vertex& Last = prd.path.last();
    st(Mid.position, geom.P);
    Mid.normal = light.normal;
    Mid.type = SPC_VTYPE_HIT_LIGHT_SOURCE;
    Mid.uv.x = geom.uvx; Mid.uv.y = geom.uvy;
    LightSample ls;
    light_reverse_sample(fr, light, Mid.uv.x, Mid.uv.y, ls);
    const float lightPdf = ls.pdf;
    const float pdf_G = absf(dot(ld(Mid.normal), ray_direction) * dot(ld(Last.normal), ray_direction)) / (t_hit * t_hit);
    if (Last.isOrigin) st(Mid.flux, ld(Last.flux) * pdf_G * ls.emission);
    else st(Mid.flux, ld(Mid.flux) * ld(Last.flux) * pdf_G * ls.emission);
    Mid.lastPosition = Last.position;
    Mid.lastNormalProjection = absf(dot(ld(Last.normal), ray_direction));
    Mid.subspaceId = (int16_t)ls.subspaceId;
    Mid.lastZoneId = Last.subspaceId;
    Mid.singlePdf = Mid.singlePdf * pdf_G / absf(dot(ld(Last.normal), ray_direction));
    Mid.pdf = Last.pdf * Mid.singlePdf;
    Mid.materialId = (int16_t)tr.light;
    Mid.depth = (int16_t)(Last.depth + 1);
    if (Mid.depth == 1) {
        Mid.RMIS_pointer = 1.0f;
        return;
    }
    spc_vertex virtual_light;
    memset(&virtual_light, 0, sizeof(virtual_light));
    virtual_light.type = SPC_VTYPE_QUAD;
    virtual_light.position = Mid.position;
    virtual_light.RMIS_pointer = 1;
    virtual_light.normal = Mid.normal;
    virtual_light.pdf = lightPdf;
    virtual_light.singlePdf = lightPdf;
    st(virtual_light.flux, ls.emission);
    virtual_light.subspaceId = Mid.subspaceId;
    virtual_light.isBrdf = 0;
    Mid.RMIS_pointer = (float)(1.0 / (double)light_hit(fr, Last, virtual_light));
}

// traceEyeSubPath / traceLightSubPath (cuProg.h:409-461): closest hit with back-face culling of
// emitter quads, then the hit/miss program of the active raygen (sutil/Scene.cpp:1648-1691)
void trace_subpath(const Frame& fr, Payload& prd, f3 o, f3 d, bool light_side) {
    Hit h;
    if (!fr.sc->closest(o, d, SCENE_EPS, 1e16f, true, h)) {
        prd.done = true;                                   // __miss__BDPTVertex, raygen.cu:699-704
        return;
    }
    const Tri& tr = fr.sc->tris[h.prim];
    if (tr.light >= 0) {
        if (light_side) prd.done = true;                   // __closesthit__lightSource_subpath, hit_program.cu:239-244
        else closesthit_eye_light(fr, prd, tr, h, d);
    } else {
        closesthit_surface(fr, prd, tr, h, d, light_side);
    }
}

// visibilityTest (cuProg.h:463-502)
bool visibility_test(const Frame& fr, f3 pos_A, f3 pos_B) {
    const f3 bias_pos = pos_B - pos_A;
    const float len = length(bias_pos);
    const f3 dir = bias_pos / len;
    return !fr.sc->occluded(pos_A, dir, SCENE_EPS, len - SCENE_EPS);
}

// binary_sample (cuProg.h:245-264): the reference's own bisect, returns l
static int binary_sample(const float* cmf, int size, uint32_t& seed, float& pmf) {
    const float index = rnd(seed) * 1.0f;
    int mid = size / 2 - 1, l = 0, r = size;
    while (r - l > 1) {
        if (index < cmf[mid]) r = mid + 1;
        else l = mid + 1;
        mid = (l + r) / 2 - 1;
    }
    pmf = l == 0 ? cmf[l] : cmf[l] - cmf[l - 1];
    return l;
}

// connectVertex_SPCBPT (raygen.cu:253-303)
float connect_mis_and_eval(const Frame& fr, const spc_vertex& a, const spc_vertex& b, f3& ans_out) {
    const f3 connectVec = ld(a.position) - ld(b.position);
    const f3 connectDir = normalize(connectVec);
    const float G = absf(dot(ld(a.normal), connectDir)) * absf(dot(ld(b.normal), connectDir)) / dot(connectVec, connectVec);
    const f3 LA_DIR = normalize(ld(a.lastPosition) - ld(a.position));
    const f3 LB_DIR = normalize(ld(b.lastPosition) - ld(b.position));
    f3 fa, fb;
    const Pbr mat_a = vertex_mat(fr, a);
    fa = bsdf_eval(mat_a, ld(a.normal), -connectDir, LA_DIR) / (mat_a.brdf ? absf(dot(ld(a.normal), connectDir)) : 1.0f);
    if (!b.isOrigin) {
        const Pbr mat_b = vertex_mat(fr, b);
        fb = bsdf_eval(mat_b, ld(b.normal), connectDir, LB_DIR) / (mat_b.brdf ? absf(dot(ld(b.normal), connectDir)) : 1.0f);
    } else {
        if (dot(ld(b.normal), -connectDir) > 0.0f) fb = mk3(0.0f);
        else fb = mk3(1.0f);
    }
    const f3 contri = ld(a.flux) * ld(b.flux) * fa * fb * G;
    const float pdf = a.pdf * b.pdf;
    const float w = (b.depth == 0 ? connection_lightSource(fr, a, b) : general_connection(fr, a, b));
    const f3 ans = contri / pdf * w;
    ans_out = invalid3(ans) ? mk3(0.0f) : ans;
    return w;
}

// ToneMap (raygen.cu:50-58) + make_color (src/cuda/helpers.h:35-67)
static inline uint8_t quantize8(float x) {
    x = clampf(x, 0.0f, 1.0f);
    return (uint8_t)std::min((unsigned)(x * 256.0f), 255u);
}
static inline float to_srgb(float c) {
    const float invGamma = 1.0f / 2.4f;
    const float powed = cm_powf(c, invGamma);
    return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f;
}
static uint32_t tonemap_pack(f3 accum) {
    const float luminance = 0.3f * accum.x + 0.6f * accum.y + 0.1f * accum.z;
    const float s = 1.0f + 1 * luminance / 1.5f;
    // `c * 1.0f / s` on a float4: operator/(float4,float) multiplies by the reciprocal (sutil/vec_math.h:720-724)
    const float inv = 1.0f / s;
    const f3 val = f3{accum.x * 1.0f * inv, accum.y * 1.0f * inv, accum.z * 1.0f * inv};
    const f3 c = f3{clampf(val.x, 0.f, 1.f), clampf(val.y, 0.f, 1.f), clampf(val.z, 0.f, 1.f)};
    return (uint32_t)quantize8(to_srgb(c.x)) | ((uint32_t)quantize8(to_srgb(c.y)) << 8) | ((uint32_t)quantize8(to_srgb(c.z)) << 16) | (255u << 24);
}

// __raygen__SPCBPT (raygen.cu:319-443)
void eye_pixel(const Frame& fr, int px, int py, int* first_prim, int* first_label) {
    const spc_params& P = fr.p;
    const unsigned W = P.width, H = P.height;
    const f3 eye = ld(P.eye), U = ld(P.U), V = ld(P.V), Wv = ld(P.W);
    const int subframe_index = (int)P.subframe_index;
    const unsigned image_index = (unsigned)py * W + (unsigned)px;
    uint32_t seed = tea(4, image_index, (uint32_t)subframe_index);
    // make_float2(rnd(seed), rnd(seed)) (raygen.cu:336): nvcc evaluates the arguments left to right
    // (verified on PTX), so x gets the first draw.  g++ -- which builds the reference-on-host
    // shim -- evaluates right to left; g_jitter_rtl reproduces that for the pinning test only.
    float jx = 0.5f, jy = 0.5f;
    if (subframe_index != 0) {
        if (g_jitter_rtl) { jy = rnd(seed); jx = rnd(seed); }
        else { jx = rnd(seed); jy = rnd(seed); }
    }
    const float dx = 2.0f * (((float)px + jx) / (float)W) - 1.0f;
    const float dy = 2.0f * (((float)py + jy) / (float)H) - 1.0f;
    f3 ray_direction = normalize(dx * U + dy * V + Wv);
    f3 ray_origin = eye;
    f3 result = mk3(0.0f);

    Payload payload;
    memset(&payload.path, 0, sizeof(payload.path));
    for (int k = 0; k < 3; k++) payload.path.v[k].type = SPC_VTYPE_QUAD;   // BDPTVertex default (BDPTVertex.h:55)
    payload.clear();
    payload.seed = seed;
    payload.ray_direction = ray_direction;
    payload.origin = ray_origin;
    {   // init_EyeSubpath (raygen.cu:216-231)
        payload.path.size++;
        spc_vertex& v = payload.path.cur();
        st(v.position, ray_origin);
        st(v.flux, mk3(1.0f));
        v.pdf = 1.0f;
        v.RMIS_pointer = 0;
        st(v.normal, ray_direction);
        v.isOrigin = 1;
        v.depth = 0;
        v.singlePdf = 1.0f;
        payload.path.next().singlePdf = 1.0f;
    }
    if (first_prim) *first_prim = -1;
    if (first_label) *first_label = -1;
    const spc_subspace_sampler& S = P.sampler;
    while (true) {
        ray_direction = payload.ray_direction;
        ray_origin = payload.origin;
        if (payload.done || payload.depth > fr.max_depth) break;
        const int begin_depth = payload.path.size;
        if (payload.depth == 0 && first_prim) {
            Hit h;
            *first_prim = fr.sc->closest(ray_origin, ray_direction, SCENE_EPS, 1e16f, true, h) ? h.prim : -1;
        }
        trace_subpath(fr, payload, ray_origin, ray_direction, false);
        if (payload.path.size == begin_depth) break;
        payload.depth += 1;
        if (payload.depth == 1 && first_label) *first_label = payload.path.cur().subspaceId;
        if (payload.path.cur().type == SPC_VTYPE_HIT_LIGHT_SOURCE) {
            // lightStraghtHit (raygen.cu:305-317)
            const spc_vertex& a = payload.path.cur();
            const f3 ans = ld(a.flux) / a.pdf / a.RMIS_pointer;
            if (!invalid3(ans)) result += ans;
            break;
        }
        const spc_vertex& eye_subpath = payload.path.cur();
        for (int it = 0; it < fr.connections; it++) {
            int light_id = 0;
            float pmf_firstStage = 1;
            if (P.subspace_info.light_tree)   // sampleFirstStage (cuProg.h:290-301)
                light_id = binary_sample(P.subspace_info.CMFGamma + (size_t)eye_subpath.subspaceId * fr.K, fr.K, payload.seed, pmf_firstStage);
            if (S.subspace[light_id].size == 0) continue;
            float pmf_secondStage;        // sampleSecondStage (cuProg.h:268-280)
            const int begin_index = S.subspace[light_id].jump_bias;
            const int index = binary_sample(S.cmfs + begin_index, S.subspace[light_id].size, payload.seed, pmf_secondStage) + begin_index;
            const spc_vertex& light_subpath = S.LVC[S.jump_buffer[index]];
            if (visibility_test(fr, ld(eye_subpath.position), ld(light_subpath.position))) {
                const float pmf = (float)S.path_count * pmf_secondStage * pmf_firstStage;
                f3 c;
                connect_mis_and_eval(fr, eye_subpath, light_subpath, c);
                const f3 res = c / pmf;
                if (!invalid3(res)) result += res / (float)fr.connections;
            }
        }
    }
    f3 accum_color = result;
    if (subframe_index > 0) {
        const float a = 1.0f / (float)(subframe_index + 1);
        const spc_float4& prev = P.accum_buffer[image_index];
        accum_color = lerp3(f3{prev.x, prev.y, prev.z}, accum_color, a);
    }
    P.accum_buffer[image_index] = spc_float4{accum_color.x, accum_color.y, accum_color.z, 1.0f};
    if (P.frame_buffer) P.frame_buffer[image_index] = tonemap_pack(accum_color);
}

// __raygen__lightTrace (raygen.cu:612-685)
void light_trace_core(const Frame& fr, int core) {
    const spc_light_trace_params& lt = fr.p.lt;
    uint32_t seed = tea(4, (uint32_t)core, (uint32_t)lt.launch_frame);
    Payload payload;
    memset(&payload.path, 0, sizeof(payload.path));
    for (int k = 0; k < 3; k++) payload.path.v[k].type = SPC_VTYPE_QUAD;
    payload.seed = seed;                       // a copy taken once: the hit-side stream (raygen.cu:628)
    const unsigned bufferBias = (unsigned)lt.core_padding * (unsigned)core;
    unsigned lightVertexCount = 0, lightPathCount = 0;
    auto push = [&](const spc_vertex& v) {     // pushVertexToLVC (raygen.cu:613-619)
        lt.ans[lightVertexCount + bufferBias] = v;
        lt.validState[lightVertexCount + bufferBias] = 1;
        lightVertexCount++;
    };
    auto full = [&]() { return !(lightVertexCount < (unsigned)lt.core_padding); };
    while (true) {
        payload.clear();
        const int light_id = pick_light(fr, seed);
        const spc_light& light = fr.sc->lights[light_id];
        LightSample ls;
        light_sample_pos(fr, light, seed, ls);
        light_trace_mode(ls, seed);
        f3 ray_direction = ls.direction;
        f3 ray_origin = ls.position;
        {   // init_lightSubPath_from_lightSample (raygen.cu:196-213)
            payload.path.size = 1;
            payload.path.next().singlePdf = ls.dir_pdf;
            init_vertex_from_light_sample(ls, payload.path.cur());
        }
        push(payload.path.cur());
        if (full()) break;
        while (true) {
            const int begin_depth = payload.path.size;
            trace_subpath(fr, payload, ray_origin, ray_direction, true);
            if (payload.path.size > begin_depth) {
                push(payload.path.cur());
                if (full()) break;
            }
            ray_direction = payload.ray_direction;
            ray_origin = payload.origin;
            if (payload.done || payload.depth > fr.max_depth) break;
            payload.depth += 1;
        }
        lightPathCount++;
        if (lightPathCount >= (unsigned)lt.M_per_core) break;
        if (full()) break;
    }
    for (unsigned i = lightVertexCount; i < (unsigned)lt.core_padding; i++) lt.validState[i + bufferBias] = 0;
}

// MyThrustOp::LVC_Process (cuda_thrust/device_thrust.cu:241-332): weight = (flux.x+flux.y+flux.z)/pdf with
// NaN/Inf -> 0 (:200-207); valid vertices bucketed by subspace in index order; per-bucket running fp32
// prefix sums in that order divided by the bucket sum; path_count = #valid depth-0 vertices.
void lvc_process(const spc_vertex* lvc, const uint8_t* valid, int n, int K, spc_subspace* subspace, float* cmfs, int* jump,
                 int* vertex_count, int* path_count) {
    std::vector<std::vector<int>> sj(K);
    std::vector<std::vector<float>> sp(K);
    std::vector<float> Qs(K, 0.f);
    int vc = 0, pc = 0;
    for (int i = 0; i < n; i++) {
        if (!valid[i]) continue;
        vc++;
        if (lvc[i].depth == 0) pc++;
        float res = (lvc[i].flux.x + lvc[i].flux.y + lvc[i].flux.z) / lvc[i].pdf;
        res = std::isinf(res) ? 0 : res;
        const float w = std::isnan(res) ? 0 : res;
        const int s = lvc[i].subspaceId;
        Qs[s] += w;
        sj[s].push_back(i);
        sp[s].push_back(w);
        if (sp[s].size() > 1) sp[s][sp[s].size() - 1] += sp[s][sp[s].size() - 2];
    }
    int acc = 0, bias = 0;
    for (int i = 0; i < K; i++) {
        subspace[i].id = i;
        subspace[i].jump_bias = bias;
        subspace[i].size = (int)sj[i].size();
        subspace[i].sum_pmf = Qs[i];
        subspace[i].Q = 0.f;
        bias += subspace[i].size;
        for (int j = 0; j < subspace[i].size; j++) {
            jump[acc] = sj[i][j];
            cmfs[acc] = sp[i][j] / subspace[i].sum_pmf;
            acc++;
        }
    }
    *vertex_count = vc;
    *path_count = pc;
}

}  // namespace orc

// ============================================================================================
// "pt": the reference's comparison integrator (unidirectional path tracing + next-event estimation + MIS):
// __raygen__pinhole (raygen.cu:71-170), __closesthit__radiance (hit_program.cu:439-552),
// __closesthit__lightsource (:148-180), __miss__constant_radiance (raygen.cu:687-696)
// ============================================================================================
namespace orc {

void pt_pixel(const Frame& fr, int px, int py) {
    const spc_params& P = fr.p;
    const unsigned W = P.width, H = P.height;
    const int subframe_index = (int)P.subframe_index;
    const unsigned image_index = (unsigned)py * W + (unsigned)px;
    uint32_t seed = tea(4, image_index, (uint32_t)subframe_index);
    float jx = 0.5f, jy = 0.5f;
    if (subframe_index != 0) {
        if (g_jitter_rtl) { jy = rnd(seed); jx = rnd(seed); }
        else { jx = rnd(seed); jy = rnd(seed); }
    }
    const float dx = 2.0f * (((float)px + jx) / (float)W) - 1.0f;
    const float dy = 2.0f * (((float)py + jy) / (float)H) - 1.0f;
    f3 ray_direction = normalize(dx * ld(P.U) + dy * ld(P.V) + ld(P.W));
    f3 ray_origin = ld(P.eye);
    // PayloadRadiance (whitted.h:86-108)
    f3 result = mk3(0.0f), throughput = mk3(1.0f), currentResult = mk3(0.0f), vis_A = mk3(0.0f), vis_B = mk3(0.0f);
    float prd_pdf = 0.f;
    int depth = 0;
    bool done = false;
    while (true) {
        Hit h;
        if (!fr.sc->closest(ray_origin, ray_direction, SCENE_EPS, 1e16f, true, h)) {
            done = true;                       // __miss__constant_radiance (no sky)
            currentResult = mk3(0.0f);
        } else {
            const Tri& tr = fr.sc->tris[h.prim];
            const LocalGeom geom = local_geometry(tr, h.u, h.v);
            if (tr.light >= 0) {               // __closesthit__lightsource
                LightSample ls;
                light_reverse_sample(fr, fr.sc->lights[tr.light], geom.uvx, geom.uvy, ls);
                const f3 ln = ld(ls.light->normal);
                if (dot(ray_direction, ln) <= 0) {
                    float MIS_weight = 1;
                    if (depth != 0) {
                        const float pdf_hit = prd_pdf * absf(dot(ray_direction, ln)) / (h.t * h.t);
                        const float pdf_area = ls.pdf;
                        MIS_weight = pdf_hit / (pdf_area + pdf_hit);
                    }
                    result += throughput * ls.emission * MIS_weight;
                }
                done = true;
            } else {                           // __closesthit__radiance
                const Pbr pbr = shade_pbr(*fr.sc, tr.material, geom.uvx, geom.uvy);
                f3 N = geom.Ng;
                if (dot(N, ray_direction) > 0.f) N = -N;
                const f3 in_dir = -ray_direction;
                f3 res = mk3(0.0f);
                const float rr_rate = std::fmax(0.3f, std::fmin(fmax3(pbr.base_color), 1.0f));   // clamp(fmaxf(color), MIN_RR_RATE, 1.0)
                const int light_id = pick_light(fr, seed);
                const spc_light& light = fr.sc->lights[light_id];
                LightSample ls;
                light_sample_pos(fr, light, seed, ls);
                const float L_dist = length(ls.position - geom.P);
                const f3 L = (ls.position - geom.P) / L_dist;
                const f3 V = -normalize(ray_direction);
                const f3 LN = ld(light.normal);
                const float L_dot_LN = dot(-L, LN);
                const float N_dot_L = dot(N, L);
                const float N_dot_V = dot(N, V);
                if (N_dot_L > 0.0f && N_dot_V > 0.0f && L_dot_LN > 0.0f) {
                    vis_A = geom.P;
                    vis_B = ls.position;
                    const f3 eval = bsdf_eval(pbr, N, V, L);
                    const float pdf_area = ls.pdf;
                    const float pdf_hit = bsdf_pdf(pbr, N, V, L) * absf(L_dot_LN) / (L_dist * L_dist) * rr_rate;
                    const float MIS_weight = pdf_area / (pdf_hit + pdf_area);
                    res += throughput * ls.emission * 1.0f / ls.pdf * N_dot_L * L_dot_LN / L_dist / L_dist * eval * MIS_weight;
                }
                currentResult += res;
                ray_origin = geom.P;
                if (rnd(seed) > rr_rate) {
                    done = true;
                } else {
                    ray_direction = bsdf_sample(pbr, N, in_dir, seed);
                    const float pdf = bsdf_pdf(pbr, N, in_dir, ray_direction);
                    if (pdf > 0.0f) {
                        throughput *= bsdf_eval(pbr, N, in_dir, ray_direction) * absf(dot(ray_direction, N)) / pdf / rr_rate;
                        prd_pdf = pdf * rr_rate;
                    } else {
                        done = true;
                    }
                }
            }
        }
        if (sum3(currentResult) > 0.0) {
            if (visibility_test(fr, vis_A, vis_B)) result += currentResult;
            currentResult = mk3(0.0f);
        }
        if (done || depth > 30) break;
        depth += 1;
    }
    f3 accum_color = result;
    if (subframe_index > 0) {
        const float a = 1.0f / (float)(subframe_index + 1);
        const spc_float4& prev = P.accum_buffer[image_index];
        accum_color = lerp3(f3{prev.x, prev.y, prev.z}, accum_color, a);
    }
    P.accum_buffer[image_index] = spc_float4{accum_color.x, accum_color.y, accum_color.z, 1.0f};
    if (P.frame_buffer) P.frame_buffer[image_index] = tonemap_pack(accum_color);
}

}  // namespace orc
