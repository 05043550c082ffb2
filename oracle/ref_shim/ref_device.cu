// ref_device.cu -- "reference programs on the GPU": the reference's OWN raygen.cu / hit_program.cu / cuProg.h / rmis.h /
// src/cuda/*.h compiled UNMODIFIED for sm_100a (nvcc --use_fast_math, the reference's own flags, src/CMakeLists.txt:214-215) from
// where they lie under /root/reference, against the device stub <optix.h> in ref_shim/stub_device/.  SURVEY.md section 8c "T1".
//
// BASELINE / TEST INFRASTRUCTURE: builds only into oracle/_ref/libref_device.so (git-ignored; travels to the GPU box like the
// other prebuilt checkers).  It is the GPU-side reference arm of bench.py (the north star's "reference build on one B200": the
// OptiX SDK is absent, so OptiX's own traversal cannot run) and a statistical checker of the fast-arithmetic flavour.  Nothing
// of the reference is copied into the repository and nothing in the product links or loads this file.
//
// What is the reference's and what is not:
//   * reference: every raygen / closest-hit / miss program (the megakernel structure included: one thread per pixel runs a whole
//     path with its connections), BSDF, light sampling, RMIS, subspace sampling, classification, RNG, tone map, with
//     --use_fast_math arithmetic and CUDA's own texture unit (9-bit filter weights) -- as the reference runs them.
//   * not reference: optixTrace.  OptiX 7.5 is closed source and B200 has no RT cores (OptiX would fall back to a software
//     traversal there); the shim's optixTrace is THIS repository's traversal (csrc/traverse.cuh, traverse_bvh8) over THIS
//     repository's BVH, called inline from the reference's megakernel, followed by a direct call of the closest-hit / miss program
//     the reference's SBT would select (sutil/Scene.cpp:1642-1691).
#include <optix.h>

#include <vector>

#include "../../spcbpt-optix7_b200/csrc/traverse.cuh"

// the reference's programs are `extern "C" __global__`: inside this translation unit they become device functions that the
// wrapper kernels below call
#undef __global__
#define __global__ __device__
#include "hit_program.cu"
#include "raygen.cu"
#undef __global__
#define __global__ __location__(global)

// nvcc's host-side registration stub names the `extern "C"` launch-parameter symbol of cuProg.h:65-67 as ::params although it is
// declared inside namespace Tracer: make that name resolve
using Tracer::params;

__constant__ int   ref_shim_kind;   // 0 pt, 1 SPCBPT_eye, 2 light trace, 3 pretrace (the four switchRaygen tables)

struct RefDevScene {
    const float4* nodes;            // this repository's BVH8 (csrc/common.cuh)
    const float4* tris;
    const float4* tri_pos;          // per prim: .w of the third vertex = mesh index
    const int*    mesh_first;       // first global prim of every mesh
    const unsigned char* mesh_is_light;
    const whitted::HitGroupData* records;   // one per mesh (sutil/Scene.cpp:1725-1771 packs one per mesh and ray type)
};
__constant__ RefDevScene ref_scene;

__device__ __forceinline__ uint2* ref_stack_base() { return reinterpret_cast<uint2*>(ref_shim_state + REF_SHIM_BLOCK); }
__device__ __forceinline__ spc::TravLut& ref_lut() { return *reinterpret_cast<spc::TravLut*>(ref_stack_base() + spc::kSmStack * REF_SHIM_BLOCK); }
constexpr size_t kRefSmem = REF_SHIM_BLOCK * sizeof(RefShimState) + spc::kSmStack * REF_SHIM_BLOCK * sizeof(uint2) + sizeof(spc::TravLut);

// optixTrace: traversal, then the program the reference's SBT would run
__device__ void ref_shim_trace(float3 o, float3 d, float tmin, float tmax, unsigned int flags, unsigned int* p0, unsigned int* p1) {
    const spc::TravRay r{o.x, o.y, o.z, d.x, d.y, d.z, tmin, tmax};
    spc::TravHit h;
    unsigned cn = 0, ct = 0;
    uint2* stack = ref_stack_base() + threadIdx.x;
    if (flags & OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT) {
        // RAY_TYPE_OCCLUSION: hit -> __closesthit__occlusion, miss -> null program (Scene.cpp:1516)
        if (spc::traverse_bvh8<true, false>(ref_scene.nodes, ref_scene.tris, r, false, stack, REF_SHIM_BLOCK, h, cn, ct, ref_lut())) {
            const RefShimState saved = REF_SHIM;
            REF_SHIM.payload[0] = *p0;
            __closesthit__occlusion();
            *p0 = REF_SHIM.payload[0];
            REF_SHIM = saved;
        }
        return;
    }
    const bool cull = (flags & OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES) != 0;
    const bool hit = spc::traverse_bvh8<false, false>(ref_scene.nodes, ref_scene.tris, r, cull, stack, REF_SHIM_BLOCK, h, cn, ct, ref_lut());
    const RefShimState saved = REF_SHIM;
    REF_SHIM.payload[0] = *p0;
    REF_SHIM.payload[1] = p1 ? *p1 : 0u;
    REF_SHIM.ray_dir = d;
    const int kind = ref_shim_kind;
    if (hit) {
        const int mesh = __float_as_int(__ldg(ref_scene.tri_pos + 3 * (size_t)h.prim + 2).w);
        REF_SHIM.sbt_data = (unsigned long long)(ref_scene.records + mesh);
        REF_SHIM.prim_index = (unsigned)(h.prim - ref_scene.mesh_first[mesh]);
        REF_SHIM.bary = make_float2(h.u, h.v);
        REF_SHIM.ray_tmax = h.t;
        const bool light = ref_scene.mesh_is_light[mesh] != 0;
        if (kind == 0) {
            if (light) __closesthit__lightsource(); else __closesthit__radiance();
        } else if (kind == 2) {
            if (light) __closesthit__lightSource_subpath(); else __closesthit__lightSubpath();
        } else {
            if (light) __closesthit__eyeSubpath_LightSource(); else __closesthit__eyeSubpath();
        }
    } else {
        if (kind == 0) __miss__constant_radiance(); else __miss__BDPTVertex();
    }
    REF_SHIM = saved;
}

// optixLaunch stand-ins: one thread per launch index, 1-D blocks of REF_SHIM_BLOCK threads
template <int KIND>
__global__ void __launch_bounds__(REF_SHIM_BLOCK) ref_launch_kernel(unsigned int total) {
    spc::trav_lut_init(ref_lut());   // ends in __syncthreads: before any thread leaves
    if (blockIdx.x * REF_SHIM_BLOCK + threadIdx.x >= total) return;
    if (KIND == 0) __raygen__pinhole();
    else if (KIND == 1) __raygen__SPCBPT();
    else if (KIND == 2) __raygen__lightTrace();
    else __raygen__TrainData();
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
#define REF_API extern "C" __attribute__((visibility("default")))
#define REF_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            snprintf(g_err, sizeof(g_err), "%s: %s", #call, cudaGetErrorString(e__)); \
            return -1;                                                              \
        }                                                                           \
    } while (0)

namespace {
char g_err[512] = "";
struct Owned {
    std::vector<void*> dev;
    std::vector<cudaArray_t> arrays;
    std::vector<cudaTextureObject_t> tex;
    Light* lights = nullptr;
    MaterialData::Pbr* pbr = nullptr;
    int n_lights = 0, n_pbr = 0;
    bool ready = false;
} g;

template <class T>
T* upload(const T* host, size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
    if (n) cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice);
    g.dev.push_back(p);
    return static_cast<T*>(p);
}
}  // namespace

REF_API const char* refdev_last_error() { return g_err; }

REF_API void refdev_scene_destroy() {
    for (void* p : g.dev) cudaFree(p);
    for (cudaTextureObject_t t : g.tex) cudaDestroyTextureObject(t);
    for (cudaArray_t a : g.arrays) cudaFreeArray(a);
    g = Owned();
}

// `ctx`: a context of THIS repository's library holding the same scene (same mesh order): its BVH and prim table are borrowed.
REF_API int refdev_scene_create(spc_context* ctx, const spc_mesh* meshes, int n_meshes, const spc_pbr* materials, int n_materials,
                                const spc_light* lights, int n_lights, const spc_texture* textures, int n_textures) {
    refdev_scene_destroy();
    if (!ctx || !ctx->c.has_scene) {
        snprintf(g_err, sizeof(g_err), "refdev_scene_create: the context has no scene");
        return -1;
    }
    static_assert(sizeof(Light) == sizeof(spc_light), "Light layout");
    static_assert(sizeof(MaterialData::Pbr) == sizeof(spc_pbr), "Pbr layout");
    static_assert(sizeof(MyParams) == sizeof(spc_params), "MyParams layout");
    REF_CUDA(cudaSetDevice(ctx->c.device));
    // textures: CUDA arrays + texture objects as sutil::Scene::addImage / addSampler make them (Scene.cpp:575-650, scene_shift.cpp:57-60)
    for (int t = 0; t < n_textures; t++) {
        cudaChannelFormatDesc cd = cudaCreateChannelDesc<uchar4>();
        cudaArray_t arr = nullptr;
        REF_CUDA(cudaMallocArray(&arr, &cd, textures[t].width, textures[t].height));
        g.arrays.push_back(arr);
        const size_t pitch = (size_t)textures[t].width * 4;
        REF_CUDA(cudaMemcpy2DToArray(arr, 0, 0, textures[t].rgba, pitch, pitch, textures[t].height, cudaMemcpyHostToDevice));
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = arr;
        cudaTextureDesc td = {};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 1;
        td.maxAnisotropy = 1;
        td.maxMipmapLevelClamp = 99;
        td.minMipmapLevelClamp = 0;
        td.mipmapFilterMode = cudaFilterModePoint;
        td.borderColor[0] = 1.0f;
        td.sRGB = 0;
        cudaTextureObject_t to = 0;
        REF_CUDA(cudaCreateTextureObject(&to, &rd, &td, nullptr));
        g.tex.push_back(to);
    }
    auto with_tex = [&](const spc_pbr& m) {
        MaterialData::Pbr p;
        memcpy((void*)&p, &m, sizeof(p));
        const uint64_t id = m.base_color_tex.tex;   // 1 + index into `textures` (include/spcbpt_b200.h), 0 = none
        p.base_color_tex.tex = (id >= 1 && id <= (uint64_t)n_textures) ? g.tex[id - 1] : 0;
        p.metallic_roughness_tex.tex = 0;
        return p;
    };
    // params.materials: all scene materials, then one per quad light (Material_shift, scene_shift.cpp:62-104)
    std::vector<MaterialData::Pbr> pbr((size_t)n_materials + n_lights);
    for (int i = 0; i < n_materials; i++) pbr[i] = with_tex(materials[i]);
    for (int i = 0; i < n_lights; i++) {
        MaterialData mtl;
        mtl.light_id = i;
        pbr[n_materials + i] = mtl.pbr;
    }
    g.pbr = upload(pbr.data(), pbr.size());
    g.n_pbr = (int)pbr.size();
    g.lights = upload(reinterpret_cast<const Light*>(lights), (size_t)n_lights);
    g.n_lights = n_lights;
    std::vector<unsigned char> rec_bytes((size_t)n_meshes * sizeof(whitted::HitGroupData), 0);   // (the struct has a union: no default constructor)
    whitted::HitGroupData* rec = reinterpret_cast<whitted::HitGroupData*>(rec_bytes.data());
    std::vector<int> first((size_t)n_meshes);
    std::vector<unsigned char> is_light((size_t)n_meshes);
    int prim = 0;
    for (int m = 0; m < n_meshes; m++) {
        const spc_mesh& me = meshes[m];
        first[m] = prim;
        prim += (int)me.n_triangles;
        float* dpos = upload(me.positions, (size_t)me.n_vertices * 3);
        uint32_t* didx = upload(me.indices, (size_t)me.n_triangles * 3);
        std::vector<float> uv((size_t)me.n_vertices * 2, 0.f);      // zero fill: scene_shift.cpp:203-206
        if (me.texcoords) memcpy(uv.data(), me.texcoords, uv.size() * sizeof(float));
        float* duv = upload(uv.data(), uv.size());
        if (!dpos || !didx || !duv) {
            snprintf(g_err, sizeof(g_err), "refdev_scene_create: out of device memory");
            return -1;
        }
        rec[m].geometry_data.type = GeometryData::TRIANGLE_MESH;
        GeometryData::TriangleMesh& tm = rec[m].geometry_data.triangle_mesh;
        tm.positions.data = (CUdeviceptr)dpos; tm.positions.count = me.n_vertices;
        tm.positions.byte_stride = 12; tm.positions.elmt_byte_size = 12;
        tm.indices.data = (CUdeviceptr)didx; tm.indices.count = 3 * me.n_triangles;
        tm.indices.byte_stride = 4; tm.indices.elmt_byte_size = 4;
        for (int j = 0; j < (int)GeometryData::num_textcoords; j++) {
            tm.texcoords[j].data = (CUdeviceptr)duv; tm.texcoords[j].count = me.n_vertices;
            tm.texcoords[j].byte_stride = 8; tm.texcoords[j].elmt_byte_size = 8;
        }
        MaterialData md;
        if (me.light_id >= 0) {
            md.emissive_factor = reinterpret_cast<const Light*>(lights)[me.light_id].quad.emission;
            md.light_id = me.light_id;
            md.id = n_materials + me.light_id;
            is_light[m] = 1;
        } else {
            md.doubleSided = true;
            md.pbr = with_tex(materials[me.material_id]);
            md.id = me.material_id;
            is_light[m] = 0;
        }
        rec[m].material_data = md;
    }
    RefDevScene s;
    s.nodes = ctx->c.bvh.nodes.p;
    s.tris = ctx->c.bvh.tris.p;
    s.tri_pos = ctx->c.geom.tri_pos.p;
    s.mesh_first = upload(first.data(), first.size());
    s.mesh_is_light = upload(is_light.data(), is_light.size());
    s.records = reinterpret_cast<const whitted::HitGroupData*>(upload(rec_bytes.data(), rec_bytes.size()));
    REF_CUDA(cudaMemcpyToSymbol(ref_scene, &s, sizeof(s)));
    REF_CUDA(cudaFuncSetAttribute(ref_launch_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRefSmem));
    REF_CUDA(cudaFuncSetAttribute(ref_launch_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRefSmem));
    REF_CUDA(cudaFuncSetAttribute(ref_launch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRefSmem));
    REF_CUDA(cudaFuncSetAttribute(ref_launch_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRefSmem));
    // the reference's programs recurse through optixTrace (trace -> closest hit -> visibility test -> trace): give them room
    size_t want = 24 * 1024, have = 0;
    cudaDeviceGetLimit(&have, cudaLimitStackSize);
    if (have < want) REF_CUDA(cudaDeviceSetLimit(cudaLimitStackSize, want));
    REF_CUDA(cudaDeviceSynchronize());
    g.ready = true;
    return 0;
}

// switchRaygen(kind) + memcpy(d_params) + optixLaunch(pipeline, stream, d_params, sizeof(MyParams), sbt, w, h, 1)
// (optixPathTracer.cpp:495-512, 527-544, 616-632).  `my_params`: the reference's own MyParams (host copy, DEVICE pointers inside);
// lights / materials are filled in from the scene as Scene::finalize does.  Asynchronous on `stream`.
REF_API int refdev_launch(const void* my_params, int kind, int w, int h, void* stream) {
    if (!g.ready) {
        snprintf(g_err, sizeof(g_err), "refdev_launch: no scene");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    MyParams p;
    memcpy((void*)&p, my_params, sizeof(p));
    p.lights.data = (CUdeviceptr)g.lights;
    p.lights.count = (unsigned)g.n_lights;
    p.lights.byte_stride = sizeof(Light);
    p.lights.elmt_byte_size = sizeof(Light);
    p.materials.data = (CUdeviceptr)g.pbr;
    p.materials.count = (unsigned)g.n_pbr;
    p.materials.byte_stride = sizeof(MaterialData::Pbr);
    p.materials.elmt_byte_size = sizeof(MaterialData::Pbr);
    REF_CUDA(cudaMemcpyToSymbolAsync(Tracer::params, &p, sizeof(p), 0, cudaMemcpyHostToDevice, st));
    const uint3 dims = make_uint3((unsigned)w, (unsigned)h, 1u);
    REF_CUDA(cudaMemcpyToSymbolAsync(ref_shim_dims, &dims, sizeof(dims), 0, cudaMemcpyHostToDevice, st));
    REF_CUDA(cudaMemcpyToSymbolAsync(ref_shim_kind, &kind, sizeof(kind), 0, cudaMemcpyHostToDevice, st));
    const unsigned total = (unsigned)w * (unsigned)h;
    const unsigned blocks = (total + REF_SHIM_BLOCK - 1) / REF_SHIM_BLOCK;
    switch (kind) {
        case 0: ref_launch_kernel<0><<<blocks, REF_SHIM_BLOCK, kRefSmem, st>>>(total); break;
        case 1: ref_launch_kernel<1><<<blocks, REF_SHIM_BLOCK, kRefSmem, st>>>(total); break;
        case 2: ref_launch_kernel<2><<<blocks, REF_SHIM_BLOCK, kRefSmem, st>>>(total); break;
        case 3: ref_launch_kernel<3><<<blocks, REF_SHIM_BLOCK, kRefSmem, st>>>(total); break;
        default: snprintf(g_err, sizeof(g_err), "refdev_launch: kind %d", kind); return -1;
    }
    REF_CUDA(cudaGetLastError());
    return 0;
}

REF_API int refdev_num_subspace() { return NUM_SUBSPACE; }
