// ref_host.cpp -- "reference on a host shim": the reference's OWN sources (raygen.cu, hit_program.cu,
// cuProg.h, rmis.h, classTree_*.h, src/cuda/*.h, optixPathTracer.h) compiled unmodified for the host
// from where they lie under /root/reference, against the stub <optix.h> in this directory.
// TEST INFRASTRUCTURE: builds only into oracle/_ref/ (git-ignored), used to pin the CPU oracle and to
// generate tests/golden/*.  Nothing of the reference is copied into the repository.
//
// What is the reference's and what is not:
//   * reference: every raygen / closest-hit / miss program, BSDF, light sampling, RMIS, subspace
//     sampling, classification tree build + lookup, RNG, tonemap, struct layouts.
//   * not reference (OptiX is closed source): ray/triangle intersection and
//     traversal = oracle/orc_scene.cpp (the intersection contract); texture filtering = the fp32
//     bilinear/wrap fetch below (CUDA's 9-bit filter weights are not reproducible); libm instead
//     of --use_fast_math intrinsics.
#include <unistd.h>
#include <optix.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

thread_local RefShimState g_shim;

// Transcendentals: the reference is built with --use_fast_math (__sinf, __powf, ...), which a host build cannot
// reproduce.  The shim gives sinf/cosf/logf/powf/expf the "contract" definition used by the oracle and by the
// product's kernels: the fp64 libm function rounded once to fp32 (oracle/orc_math.h cm_*).  All standard
// headers the reference pulls in are included above, before these macros.
#include <fstream>
#include <iostream>
#include <sstream>
#include <map>
#include <set>
#include <queue>
#include <random>
static inline float ref_cm_sinf(float x) { return (float)::sin((double)x); }
static inline float ref_cm_cosf(float x) { return (float)::cos((double)x); }
static inline float ref_cm_logf(float x) { return (float)::log((double)x); }
static inline float ref_cm_expf(float x) { return (float)::exp((double)x); }
static inline float ref_cm_powf(float x, float y) { return (float)::pow((double)x, (double)y); }
#define sinf ref_cm_sinf
#define cosf ref_cm_cosf
#define logf ref_cm_logf
#define expf ref_cm_expf
#define powf ref_cm_powf

#ifdef REF_VARIANT_HEADER
// Variant build (oracle/Makefile, _ref/libref_host_k64c2d6.so): the reference's optixPathTracer.h with other values for its compile-time
// constants (NUM_SUBSPACE, NUM_SUBSPACE_LIGHTSOURCE, CONNECTION_N), pre-included so that the original (same include guard) becomes a
// no-op; raygen.cu comes from the variant directory too (its depth limit `> 50` replaced).  Pins the oracle's RUNTIME K / connections /
// max_depth against the reference's own code compiled for those values.
#include REF_VARIANT_HEADER
#endif
#include "rmis_patched.h"   // see oracle/Makefile: rmis.h with getMat() returning by value
#include "hit_program.cu"
#include "raygen.cu"
#include "decisionTree/classTree_host.h"

#include "../../include/spcbpt_b200.h"
#include "../orc_scene.h"

namespace {

struct RefScene {
    orc::Scene* geo = nullptr;                       // intersection only
    std::vector<unsigned char> record_bytes;         // one whitted::HitGroupData per mesh (radiance ray type)
    whitted::HitGroupData* record(int m) { return reinterpret_cast<whitted::HitGroupData*>(record_bytes.data()) + m; }
    std::vector<int> prim_mesh, prim_local;          // global prim -> (mesh, local prim)
    std::vector<bool> mesh_is_light;
    std::vector<MaterialData::Pbr> pbr;              // params.materials
    std::vector<Light> lights;                       // params.lights
    std::vector<std::vector<float3>> pos;
    std::vector<std::vector<unsigned int>> idx;
    std::vector<std::vector<Vec2f>> uv;
    std::vector<orc::Texture> textures;
};
RefScene* g_scene = nullptr;
int g_kind = 1;   // 0 pt, 1 SPCBPT_eye, 2 light trace, 3 pretrace

}  // namespace

// fp32 bilinear, wrap addressing, normalized coordinates, texel centres at +0.5 (the addressing of
// cudaFilterModeLinear / cudaAddressModeWrap, scene_shift.cpp:57-60) on RGBA8 -> [0,1] floats.
float4 ref_shim_tex2D(cudaTextureObject_t tex, float u, float v) {
    const orc::Texture& t = g_scene->textures[(size_t)tex - 1];
    const float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    auto wrap = [](int i, int n) { i %= n; return i < 0 ? i + n : i; };
    const int x0 = wrap((int)fx, t.w), x1 = wrap((int)fx + 1, t.w), y0 = wrap((int)fy, t.h), y1 = wrap((int)fy + 1, t.h);
    float out[4];
    for (int c = 0; c < 4; c++) {
        const float t00 = t.rgba[4 * ((size_t)y0 * t.w + x0) + c] * (1.0f / 255.0f);
        const float t10 = t.rgba[4 * ((size_t)y0 * t.w + x1) + c] * (1.0f / 255.0f);
        const float t01 = t.rgba[4 * ((size_t)y1 * t.w + x0) + c] * (1.0f / 255.0f);
        const float t11 = t.rgba[4 * ((size_t)y1 * t.w + x1) + c] * (1.0f / 255.0f);
        const float a = t00 + ax * (t10 - t00), b = t01 + ax * (t11 - t01);
        out[c] = a + ay * (b - a);
    }
    return make_float4(out[0], out[1], out[2], out[3]);
}

// optixTrace: closest hit / occlusion by the oracle's intersector, then the program the reference's
// SBT would run (sutil/Scene.cpp:1648-1691 switchRaygen table, :1725-1771 record packing).
void ref_shim_trace(float3 o, float3 d, float tmin, float tmax, unsigned int flags, unsigned int sbt_offset,
                    unsigned int* p0, unsigned int* p1) {
    const RefShimState saved = g_shim;
    g_shim.payload[0] = *p0;
    g_shim.payload[1] = p1 ? *p1 : 0u;
    const orc::f3 oo = orc::mk3(o.x, o.y, o.z), dd = orc::mk3(d.x, d.y, d.z);
    if (flags & OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT) {
        // RAY_TYPE_OCCLUSION: hit -> __closesthit__occlusion, miss -> null program (Scene.cpp:1516)
        if (g_scene->geo->occluded(oo, dd, tmin, tmax)) __closesthit__occlusion();
        *p0 = g_shim.payload[0];
    } else {
        orc::Hit h;
        const bool cull = (flags & OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES) != 0;
        if (g_scene->geo->closest(oo, dd, tmin, tmax, cull, h)) {
            const int mesh = g_scene->prim_mesh[h.prim];
            g_shim.sbt_data = g_scene->record(mesh);
            g_shim.prim_index = (unsigned)g_scene->prim_local[h.prim];
            g_shim.bary = make_float2(h.u, h.v);
            g_shim.ray_tmax = h.t;
            g_shim.ray_dir = d;
            const bool light = g_scene->mesh_is_light[mesh];
            switch (g_kind) {
                case 0: light ? __closesthit__lightsource() : __closesthit__radiance(); break;
                case 2: light ? __closesthit__lightSource_subpath() : __closesthit__lightSubpath(); break;
                default: light ? __closesthit__eyeSubpath_LightSource() : __closesthit__eyeSubpath(); break;
            }
        } else {
            g_shim.ray_dir = d;
            if (g_kind == 0) __miss__constant_radiance(); else __miss__BDPTVertex();
        }
    }
    const unsigned int keep0 = g_shim.payload[0];
    g_shim = saved;
    (void)keep0;
    (void)sbt_offset;
}

#define REF_API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------------
// T0: layout / RNG / classification tree, straight from the reference headers
// ---------------------------------------------------------------------------------------------
REF_API const char* ref_layout_json() {
    static std::string s;
    char buf[8192];
#define OFF(T, f) (int)offsetof(T, f)
    snprintf(buf, sizeof(buf),
        "{\"sizeof\": {\"BDPTVertex\": %zu, \"MyParams\": %zu, \"Light\": %zu, \"MaterialData::Pbr\": %zu, \"MaterialData\": %zu, "
        "\"tree_node\": %zu, \"Subspace\": %zu, \"divide_weight\": %zu, \"pathInfo_node\": %zu, \"pathInfo_sample\": %zu, "
        "\"nVertex\": %zu, \"HitGroupData\": %zu, \"LightTraceParams\": %zu, \"PreTraceParams\": %zu, \"SubspaceSampler\": %zu, "
        "\"subspaceMacroInfo\": %zu, \"envInfo\": %zu, \"BDPTPath\": %zu},"
        " \"train\": {\"sample.sample_pdf\": %d, \"sample.fix_pdf\": %d, \"sample.begin_ind\": %d, \"sample.end_ind\": %d, \"sample.choice_id\": %d, "
        "\"sample.pixel_id\": %d, \"sample.valid\": %d, \"node.B_position\": %d, \"node.A_dir_d\": %d, \"node.B_dir_d\": %d, \"node.A_normal_d\": %d, "
        "\"node.B_normal_d\": %d, \"node.peak_pdf\": %d, \"node.path_id\": %d, \"node.label_A\": %d, \"node.label_B\": %d, \"node.valid\": %d, \"node.light_source\": %d},"
        " \"offsetof\": {\"MyParams\": {\"width\": %d, \"height\": %d, \"subframe_index\": %d, \"accum_buffer\": %d, \"frame_buffer\": %d, "
        "\"max_depth\": %d, \"eye\": %d, \"U\": %d, \"V\": %d, \"W\": %d, \"lights\": %d, \"materials\": %d, \"miss_color\": %d, "
        "\"handle\": %d, \"lt\": %d, \"sampler\": %d, \"pre_tracer\": %d, \"subspace_info\": %d, \"sky\": %d},"
        " \"BDPTVertex\": {\"position\": %d, \"normal\": %d, \"flux\": %d, \"color\": %d, \"lastPosition\": %d, \"RMIS_pointer_3\": %d, "
        "\"uv\": %d, \"RMIS_pointer\": %d, \"last_lum\": %d, \"lastNormalProjection\": %d, \"pdf\": %d, \"singlePdf\": %d, "
        "\"lastSinglePdf\": %d, \"materialId\": %d, \"subspaceId\": %d, \"depth\": %d, \"lastZoneId\": %d, \"type\": %d, "
        "\"isOrigin\": %d, \"inBrdf\": %d, \"lastBrdf\": %d, \"isBrdf\": %d, \"isLastVertex_direction\": %d},"
        " \"Light\": {\"type\": %d, \"id\": %d, \"divLevel\": %d, \"ssBase\": %d, \"corner\": %d, \"u\": %d, \"v\": %d, "
        "\"emission\": %d, \"normal\": %d, \"area\": %d},"
        " \"Pbr\": {\"base_color\": %d, \"metallic\": %d, \"roughness\": %d, \"specular\": %d, \"specularTint\": %d, \"subsurface\": %d, "
        "\"anisotropic\": %d, \"sheen\": %d, \"sheenTint\": %d, \"clearcoat\": %d, \"clearcoatGloss\": %d, \"base_color_tex\": %d, "
        "\"metallic_roughness_tex\": %d, \"brdf\": %d}},"
        " \"constants\": {\"NUM_SUBSPACE\": %d, \"NUM_SUBSPACE_LIGHTSOURCE\": %d, \"CONNECTION_N\": %d, \"MIN_RR_RATE\": %g, "
        "\"CONSERVATIVE_RATE\": %g, \"DIR_JUDGE\": %d, \"PRETRACE_CONN_PADDING\": %d, \"SCENE_EPSILON\": %g}}",
        sizeof(BDPTVertex), sizeof(MyParams), sizeof(Light), sizeof(MaterialData::Pbr), sizeof(MaterialData),
        sizeof(classTree::tree_node), sizeof(Subspace), sizeof(classTree::divide_weight), sizeof(TrainData::pathInfo_node),
        sizeof(TrainData::pathInfo_sample), sizeof(TrainData::nVertex), sizeof(whitted::HitGroupData), sizeof(LightTraceParams),
        sizeof(PreTraceParams), sizeof(SubspaceSampler), sizeof(subspaceMacroInfo), sizeof(envInfo), sizeof(BDPTPath),
        OFF(TrainData::pathInfo_sample, sample_pdf), OFF(TrainData::pathInfo_sample, fix_pdf), OFF(TrainData::pathInfo_sample, begin_ind),
        OFF(TrainData::pathInfo_sample, end_ind), OFF(TrainData::pathInfo_sample, choice_id), OFF(TrainData::pathInfo_sample, pixel_id),
        OFF(TrainData::pathInfo_sample, valid), OFF(TrainData::pathInfo_node, B_position), OFF(TrainData::pathInfo_node, A_dir_d),
        OFF(TrainData::pathInfo_node, B_dir_d), OFF(TrainData::pathInfo_node, A_normal_d), OFF(TrainData::pathInfo_node, B_normal_d),
        OFF(TrainData::pathInfo_node, peak_pdf), OFF(TrainData::pathInfo_node, path_id), OFF(TrainData::pathInfo_node, label_A),
        OFF(TrainData::pathInfo_node, label_B), OFF(TrainData::pathInfo_node, valid), OFF(TrainData::pathInfo_node, light_source),
        OFF(MyParams, width), OFF(MyParams, height), OFF(MyParams, subframe_index), OFF(MyParams, accum_buffer), OFF(MyParams, frame_buffer),
        OFF(MyParams, max_depth), OFF(MyParams, eye), OFF(MyParams, U), OFF(MyParams, V), OFF(MyParams, W), OFF(MyParams, lights),
        OFF(MyParams, materials), OFF(MyParams, miss_color), OFF(MyParams, handle), OFF(MyParams, lt), OFF(MyParams, sampler),
        OFF(MyParams, pre_tracer), OFF(MyParams, subspace_info), OFF(MyParams, sky),
        OFF(BDPTVertex, position), OFF(BDPTVertex, normal), OFF(BDPTVertex, flux), OFF(BDPTVertex, color), OFF(BDPTVertex, lastPosition),
        OFF(BDPTVertex, RMIS_pointer_3), OFF(BDPTVertex, uv), OFF(BDPTVertex, RMIS_pointer), OFF(BDPTVertex, last_lum),
        OFF(BDPTVertex, lastNormalProjection), OFF(BDPTVertex, pdf), OFF(BDPTVertex, singlePdf), OFF(BDPTVertex, lastSinglePdf),
        OFF(BDPTVertex, materialId), OFF(BDPTVertex, subspaceId), OFF(BDPTVertex, depth), OFF(BDPTVertex, lastZoneId), OFF(BDPTVertex, type),
        OFF(BDPTVertex, isOrigin), OFF(BDPTVertex, inBrdf), OFF(BDPTVertex, lastBrdf), OFF(BDPTVertex, isBrdf), OFF(BDPTVertex, isLastVertex_direction),
        OFF(Light, type), OFF(Light, id), OFF(Light, divLevel), OFF(Light, ssBase), OFF(Light, quad.corner), OFF(Light, quad.u), OFF(Light, quad.v),
        OFF(Light, quad.emission), OFF(Light, quad.normal), OFF(Light, quad.area),
        OFF(MaterialData::Pbr, base_color), OFF(MaterialData::Pbr, metallic), OFF(MaterialData::Pbr, roughness), OFF(MaterialData::Pbr, specular),
        OFF(MaterialData::Pbr, specularTint), OFF(MaterialData::Pbr, subsurface), OFF(MaterialData::Pbr, anisotropic), OFF(MaterialData::Pbr, sheen),
        OFF(MaterialData::Pbr, sheenTint), OFF(MaterialData::Pbr, clearcoat), OFF(MaterialData::Pbr, clearcoatGloss),
        OFF(MaterialData::Pbr, base_color_tex), OFF(MaterialData::Pbr, metallic_roughness_tex), OFF(MaterialData::Pbr, brdf),
        (int)NUM_SUBSPACE, (int)NUM_SUBSPACE_LIGHTSOURCE, (int)CONNECTION_N, (double)MIN_RR_RATE, (double)CONSERVATIVE_RATE, (int)DIR_JUDGE,
        (int)PRETRACE_CONN_PADDING, (double)SCENE_EPSILON);
    s = buf;
    return s.c_str();
}

REF_API unsigned int ref_tea4(unsigned int a, unsigned int b) { return tea<4>(a, b); }
REF_API unsigned int ref_tea16(unsigned int a, unsigned int b) { return tea<16>(a, b); }
REF_API void ref_rnd_stream(unsigned int* state, int n, float* out) {
    for (int i = 0; i < n; i++) out[i] = rnd(*state);
}

// classTree::buildTreeBaseOnExistSample()(samples, K, labelBias)  (classTree_host.h:302)
REF_API int ref_tree_build(const spc_divide_weight* samples, int n, int K, int label_bias, spc_tree_node* out, int out_cap, int* max_label) {
    std::vector<classTree::divide_weight> v(n);
    static_assert(sizeof(classTree::divide_weight) == sizeof(spc_divide_weight), "divide_weight layout");
    static_assert(sizeof(classTree::tree_node) == sizeof(spc_tree_node), "tree_node layout");
    memcpy((void*)v.data(), samples, n * sizeof(spc_divide_weight));
    classTree::tree t = classTree::buildTreeBaseOnExistSample()(v, K, label_bias);
    if (max_label) *max_label = t.max_label;
    const int size = t.size;
    if (size <= out_cap) memcpy((void*)out, t.v, size * sizeof(spc_tree_node));
    delete[] t.v;
    delete[] t.center;
    return size;
}
// sutil::Camera::UVWFrame (sutil/Camera.cpp:32-43), compiled from the reference's own Camera.cpp: the camera frame the host hands to the
// raygen program (optixPathTracer.cpp:371-380 updateState / handleCameraUpdate).  out = U[3], V[3], W[3]
#include <sutil/Camera.cpp>
REF_API void ref_camera_uvw(const float* eye, const float* lookat, const float* up, float fov_y, float aspect, float* out) {
    const sutil::Camera cam(make_float3(eye[0], eye[1], eye[2]), make_float3(lookat[0], lookat[1], lookat[2]), make_float3(up[0], up[1], up[2]), fov_y, aspect);
    float3 U, V, W;
    cam.UVWFrame(U, V, W);
    const float3 v[3] = {U, V, W};
    for (int k = 0; k < 3; k++) { out[3 * k] = v[k].x; out[3 * k + 1] = v[k].y; out[3 * k + 2] = v[k].z; }
}

// StaticWorkDistribution (sutil/WorkDistribution.h:34-91), the multi-GPU tile layout the reference ships: owner[y * w + x] = the GPU whose
// getSamplePixel enumeration contains the pixel.  Returns the number of pixels enumerated twice (0 for a partition); pixels no GPU
// enumerates keep -1.
#include <sutil/WorkDistribution.h>
REF_API int ref_tile_owner_map(int w, int h, int num_gpus, int* owner) {
    for (int i = 0; i < w * h; i++) owner[i] = -1;
    int twice = 0;
    for (int g = 0; g < num_gpus; g++) {
        StaticWorkDistribution d;
        d.setRasterSize(w, h);
        d.setNumGPUs(num_gpus);
        const int n = d.numSamples(g);
        for (int s = 0; s < n; s++) {
            const int2 p = d.getSamplePixel(g, s);
            if (p.x < 0 || p.y < 0 || p.x >= w || p.y >= h) continue;
            if (owner[p.y * w + p.x] != -1) twice++;
            owner[p.y * w + p.x] = g;
        }
    }
    return twice;
}

// classTree::tree_load(eye, light) (classTree_host.h:15-60): reads tree_eye.txt / tree_light.txt from the working directory.  Run in `dir`
// to check the files host/train_state.cpp writes.  Returns 0, or -1 when a tree exceeds `cap` / the directory cannot be entered.
REF_API int ref_tree_load(const char* dir, spc_tree_node* eye, int* n_eye, spc_tree_node* light, int* n_light, int cap) {
    char old[4096];
    if (!getcwd(old, sizeof(old)) || chdir(dir) != 0) return -1;
    std::vector<classTree::tree_node> e, l;
    classTree::tree_load(e, l);
    const int back = chdir(old);
    (void)back;
    *n_eye = (int)e.size();
    *n_light = (int)l.size();
    if ((int)e.size() > cap || (int)l.size() > cap) return -1;
    if (!e.empty()) memcpy((void*)eye, e.data(), e.size() * sizeof(spc_tree_node));
    if (!l.empty()) memcpy((void*)light, l.data(), l.size() * sizeof(spc_tree_node));
    return 0;
}
// classTree::tree_index (classTree_common.h:39-51) == classTree::getLabel (classTree_device.h:8-11)
REF_API void ref_tree_index(const spc_tree_node* tree, const float* pos, const float* nrm, int n, int* labels) {
    classTree::tree_node* root = (classTree::tree_node*)tree;
    for (int i = 0; i < n; i++)
        labels[i] = classTree::tree_index(root, make_float3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]),
                                          make_float3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]), make_float3(0.0f));
}

// ---------------------------------------------------------------------------------------------
// BSDF: Tracer::Eval / Sample / Pdf (cuProg.h:735,826,868) on caller-provided inputs
// ---------------------------------------------------------------------------------------------
REF_API void ref_bsdf(const spc_pbr* mat, const float* N, const float* V, const float* L, unsigned int* seed,
                      float* eval3, float* pdf1, float* sample3) {
    MaterialData::Pbr m;
    static_assert(sizeof(MaterialData::Pbr) == sizeof(spc_pbr), "Pbr layout");
    memcpy((void*)&m, mat, sizeof(m));
    const float3 n = make_float3(N[0], N[1], N[2]), v = make_float3(V[0], V[1], V[2]), l = make_float3(L[0], L[1], L[2]);
    const float3 e = Tracer::Eval(m, n, v, l);
    eval3[0] = e.x; eval3[1] = e.y; eval3[2] = e.z;
    *pdf1 = Tracer::Pdf(m, n, v, l);
    const float3 s = Tracer::Sample(m, n, v, *seed);
    sample3[0] = s.x; sample3[1] = s.y; sample3[2] = s.z;
}

// ---------------------------------------------------------------------------------------------
// T1: the reference's programs over a host scene
// ---------------------------------------------------------------------------------------------
REF_API void ref_scene_destroy() {
    if (g_scene) {
        delete g_scene->geo;
        delete g_scene;
        g_scene = nullptr;
    }
}

REF_API int ref_scene_create(const spc_mesh* meshes, int n_meshes, const spc_pbr* materials, int n_materials,
                             const spc_light* lights, int n_lights, const spc_texture* textures, int n_textures) {
    ref_scene_destroy();
    RefScene* s = new RefScene();
    s->geo = new orc::Scene();
    static_assert(sizeof(Light) == sizeof(spc_light), "Light layout");
    s->lights.resize(n_lights);
    if (n_lights) memcpy((void*)s->lights.data(), lights, n_lights * sizeof(Light));
    // params.materials: all scene materials, then one per quad light (Material_shift, scene_shift.cpp:62-104)
    s->pbr.resize(n_materials + n_lights);
    if (n_materials) memcpy((void*)s->pbr.data(), materials, n_materials * sizeof(spc_pbr));
    for (int i = 0; i < n_lights; i++) {
        MaterialData mtl;
        mtl.light_id = i;
        s->pbr[n_materials + i] = mtl.pbr;
    }
    s->pos.resize(n_meshes); s->idx.resize(n_meshes); s->uv.resize(n_meshes);
    s->record_bytes.assign((size_t)n_meshes * sizeof(whitted::HitGroupData) + 16, 0);
    s->mesh_is_light.resize(n_meshes);
    for (int m = 0; m < n_meshes; m++) {
        const spc_mesh& me = meshes[m];
        s->pos[m].resize(me.n_vertices);
        memcpy((void*)s->pos[m].data(), me.positions, me.n_vertices * 12);
        s->idx[m].assign(me.indices, me.indices + 3 * (size_t)me.n_triangles);
        s->uv[m].resize(me.n_vertices);
        for (uint32_t v = 0; v < me.n_vertices; v++) {
            s->uv[m][v].x = me.texcoords ? me.texcoords[2 * v] : 0.f;       // zero fill: scene_shift.cpp:203-206
            s->uv[m][v].y = me.texcoords ? me.texcoords[2 * v + 1] : 0.f;
        }
        whitted::HitGroupData& rec = *s->record(m);
        rec.geometry_data.type = GeometryData::TRIANGLE_MESH;
        GeometryData::TriangleMesh& tm = rec.geometry_data.triangle_mesh;
        tm.positions.data = (CUdeviceptr)s->pos[m].data(); tm.positions.count = me.n_vertices;
        tm.positions.byte_stride = 12; tm.positions.elmt_byte_size = 12;
        tm.indices.data = (CUdeviceptr)s->idx[m].data(); tm.indices.count = 3 * me.n_triangles;
        tm.indices.byte_stride = 4; tm.indices.elmt_byte_size = 4;
        for (int j = 0; j < (int)GeometryData::num_textcoords; j++) {
            tm.texcoords[j].data = (CUdeviceptr)s->uv[m].data(); tm.texcoords[j].count = me.n_vertices;
            tm.texcoords[j].byte_stride = 8; tm.texcoords[j].elmt_byte_size = 8;
        }
        MaterialData md;
        if (me.light_id >= 0) {
            md.emissive_factor = s->lights[me.light_id].quad.emission;
            md.light_id = me.light_id;
            md.id = n_materials + me.light_id;
            s->mesh_is_light[m] = true;
        } else {
            md.doubleSided = true;
            memcpy((void*)&md.pbr, &materials[me.material_id], sizeof(spc_pbr));
            md.id = me.material_id;
            s->mesh_is_light[m] = false;
        }
        rec.material_data = md;
        for (uint32_t t = 0; t < me.n_triangles; t++) {
            orc::Tri tr;
            orc::f3 p[3];
            for (int k = 0; k < 3; k++) {
                const uint32_t vi = me.indices[3 * (size_t)t + k];
                p[k] = orc::mk3(me.positions[3 * (size_t)vi], me.positions[3 * (size_t)vi + 1], me.positions[3 * (size_t)vi + 2]);
                tr.uv[k][0] = s->uv[m][vi].x; tr.uv[k][1] = s->uv[m][vi].y;
            }
            tr.v0 = p[0]; tr.v1 = p[1]; tr.v2 = p[2];
            tr.e1 = p[1] - p[0]; tr.e2 = p[2] - p[0];
            tr.material = me.light_id >= 0 ? -1 : me.material_id;
            tr.light = me.light_id;
            tr.mesh = m;
            s->geo->tris.push_back(tr);
            s->prim_mesh.push_back(m);
            s->prim_local.push_back((int)t);
        }
    }
    for (int t = 0; t < n_textures; t++) {
        orc::Texture tx;
        tx.w = textures[t].width; tx.h = textures[t].height;
        tx.rgba.assign(textures[t].rgba, textures[t].rgba + (size_t)tx.w * tx.h * 4);
        s->textures.push_back(std::move(tx));
    }
    s->geo->build_bvh();
    g_scene = s;
    return 0;
}

// optixLaunch(pipeline, 0, d_params, sizeof(MyParams), sbt, w, h, 1) after switchRaygen(kind)
// (optixPathTracer.cpp:502-512, 534-544, 612-632).  `my_params` is the reference's own MyParams with
// HOST pointers; lights/materials/handle are filled in here from the scene.
REF_API int ref_launch(const void* my_params, int kind, int w, int h, int threads) {
    if (!g_scene) return -1;
    static_assert(sizeof(MyParams) == sizeof(spc_params), "MyParams layout");
    memcpy((void*)&Tracer::params, my_params, sizeof(MyParams));
    Tracer::params.lights.data = (CUdeviceptr)g_scene->lights.data();
    Tracer::params.lights.count = (unsigned)g_scene->lights.size();
    Tracer::params.lights.byte_stride = sizeof(Light);
    Tracer::params.lights.elmt_byte_size = sizeof(Light);
    Tracer::params.materials.data = (CUdeviceptr)g_scene->pbr.data();
    Tracer::params.materials.count = (unsigned)g_scene->pbr.size();
    Tracer::params.materials.byte_stride = sizeof(MaterialData::Pbr);
    Tracer::params.materials.elmt_byte_size = sizeof(MaterialData::Pbr);
    g_kind = kind;
    const long total = (long)w * h;
    std::atomic<long> next(0);
    auto worker = [&]() {
        for (;;) {
            const long b = next.fetch_add(256);
            if (b >= total) break;
            for (long i = b; i < std::min(total, b + 256); i++) {
                g_shim.launch_index = make_uint3((unsigned)(i % w), (unsigned)(i / w), 0u);
                g_shim.launch_dims = make_uint3((unsigned)w, (unsigned)h, 1u);
                switch (kind) {
                    case 0: __raygen__pinhole(); break;
                    case 1: __raygen__SPCBPT(); break;
                    case 2: __raygen__lightTrace(); break;
                    case 3: __raygen__TrainData(); break;
                    default: break;
                }
            }
        }
    };
    if (threads <= 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
    return 0;
}
