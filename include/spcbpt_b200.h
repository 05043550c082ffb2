/*
 * spcbpt_b200.h -- C ABI of libspcbpt_b200.so, the B200-native (sm_100a) SPCBPT render core.
 *
 * This header is the drop-in boundary for the render path of ssufujia/SPCBPT-OptiX7.  Every entry
 * point names the reference interface it stands in for (paths relative to the reference tree,
 * src/OptiXPathTracer/ unless stated otherwise).  Plain C: pointers, sizes and POD structs only.
 *
 * The POD structs below keep the reference's byte layout (sizes/offsets are static_assert-ed in
 * csrc/layout_check.cu and tested in tests/test_layout.py) so a reference host can pass its own
 * `MyParams`, `BDPTVertex`, `Light`, `MaterialData::Pbr`, `classTree::tree_node` memory unchanged.
 *
 * Error model: every call returns 0 on success or a negative spc_status; spc_last_error() returns
 * a human-readable message for the calling thread's last failure (the reference throws
 * sutil::Exception from CUDA_CHECK / OPTIX_CHECK, sutil/Exception.h:82-115).  There is no CPU
 * fallback: without a CUDA device spc_create fails with SPC_ERR_NO_DEVICE.
 */
#ifndef SPCBPT_B200_H
#define SPCBPT_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SPC_API __attribute__((visibility("default")))
#else
#define SPC_API
#endif

/* ------------------------------------------------------------------------------------------ */
/* status codes                                                                                */
/* ------------------------------------------------------------------------------------------ */
typedef enum spc_status {
    SPC_OK              = 0,
    SPC_ERR_INVALID     = -1, /* bad argument / call order                                      */
    SPC_ERR_CUDA        = -2, /* a CUDA runtime call failed (message has the CUDA error string) */
    SPC_ERR_NO_DEVICE   = -3, /* no CUDA device: there is deliberately no CPU fallback          */
    SPC_ERR_NO_SCENE    = -4, /* launch/trace before spc_scene_upload                           */
    SPC_ERR_CAPACITY    = -5  /* a caller-provided buffer is too small                          */
} spc_status;

/* ------------------------------------------------------------------------------------------ */
/* POD mirrors of the reference's host<->device contract (layout kept)   */
/* ------------------------------------------------------------------------------------------ */
typedef struct spc_float2 { float x, y; } spc_float2;          /* NB: CUDA float2 is 8-aligned  */
typedef struct spc_float3 { float x, y, z; } spc_float3;
typedef struct spc_float4 { float x, y, z, w; } spc_float4;
typedef struct spc_int2   { int32_t x, y; } spc_int2;

/* src/cuda/MaterialData.h:67-76 (MaterialData::Texture, 40 B).  `tex` is a cudaTextureObject_t in
 * the reference; here it is 1 + index into the texture table given to spc_scene_upload (0 = none):
 * textures are sampled by our own bilinear/wrap fetch so that host oracle and GPU agree. */
typedef struct spc_texture_ref {
    int32_t    texcoord;
    int32_t    _pad0;
    uint64_t   tex;
    float      texcoord_offset[2];
    float      texcoord_rotation[2];   /* sin, cos */
    float      texcoord_scale[2];
} spc_texture_ref;

/* src/cuda/MaterialData.h:78-97 (MaterialData::Pbr, 144 B, 16-aligned) */
typedef struct
#if defined(__GNUC__)
__attribute__((aligned(16)))
#endif
spc_pbr {
    float           base_color[4];
    float           metallic;
    float           roughness;
    float           specular;
    float           specularTint;
    float           subsurface;
    float           anisotropic;
    float           sheen;
    float           sheenTint;
    float           clearcoat;
    float           clearcoatGloss;
    spc_texture_ref base_color_tex;
    spc_texture_ref metallic_roughness_tex;
    uint8_t         brdf;               /* "glass" flag; unfinished in the reference (readme.md:28) */
    uint8_t         _pad1[7];
} spc_pbr;

/* src/cuda/Light.h:30-92 (Light, 80 B).  Only QUAD lights are live on the SPCBPT path. */
enum { SPC_LIGHT_POINT = 0, SPC_LIGHT_AMBIENT = 1, SPC_LIGHT_QUAD = 2, SPC_LIGHT_DIRECTIONAL = 3, SPC_LIGHT_ENV = 4 };
typedef struct spc_light {
    int32_t    type;
    int32_t    id;
    int32_t    divLevel;
    int32_t    ssBase;
    /* union member `quad` (Light.h:70-78); note u and v are corner+u, corner+v (scene_shift.cpp:127-129) */
    spc_float3 corner;
    spc_float3 u;
    spc_float3 v;
    spc_float3 emission;
    spc_float3 normal;
    float      area;
} spc_light;

/* BDPTVertex.h:9-70 (BDPTVertex, 120 B, 8-aligned) */
enum { SPC_VTYPE_SPHERE = 0, SPC_VTYPE_QUAD = 1, SPC_VTYPE_DIRECTION = 2, SPC_VTYPE_ENV = 3,
       SPC_VTYPE_HIT_LIGHT_SOURCE = 4, SPC_VTYPE_ENV_MISS = 5, SPC_VTYPE_NORMALHIT = 6 };  /* light_parameters.h:8-11 LightType */
typedef struct
#if defined(__GNUC__)
__attribute__((aligned(8)))
#endif
spc_vertex {
    spc_float3 position;
    spc_float3 normal;
    spc_float3 flux;
    spc_float3 color;
    spc_float3 lastPosition;
    spc_float3 RMIS_pointer_3;
    spc_float2 uv;
    float      RMIS_pointer;
    float      last_lum;
    float      lastNormalProjection;
    float      pdf;
    float      singlePdf;
    float      lastSinglePdf;
    int16_t    materialId;
    int16_t    subspaceId;
    int16_t    depth;
    int16_t    lastZoneId;
    int16_t    type;
    uint8_t    isOrigin;
    uint8_t    inBrdf;
    uint8_t    lastBrdf;
    uint8_t    isBrdf;
    uint8_t    isLastVertex_direction;
    uint8_t    _pad;
} spc_vertex;

/* decisionTree/classTree_common.h:11-38 (classTree::tree_node, 56 B) */
typedef struct spc_tree_node {
    spc_float3 mid;
    int32_t    child[8];
    int32_t    label;
    int32_t    type;      /* 0 position, 1 normal, 2 direction (unused: DIR_JUDGE 0) */
    uint8_t    leaf;
    uint8_t    _pad[3];
} spc_tree_node;

/* decisionTree/classTree_common.h:77-91 (classTree::divide_weight, 40 B) */
typedef struct spc_divide_weight {
    spc_float3 position;
    spc_float3 dir;
    spc_float3 normal;
    float      weight;
} spc_divide_weight;

/* optixPathTracer.h:43-51 (Subspace, 20 B) */
typedef struct spc_subspace {
    int32_t jump_bias;
    int32_t id;
    int32_t size;
    float   sum_pmf;
    float   Q;
} spc_subspace;

/* src/cuda/BufferView.h:35-63 (BufferView<T>, 16 B) */
typedef struct spc_buffer_view {
    uint64_t data;
    uint32_t count;
    uint16_t byte_stride;
    uint16_t elmt_byte_size;
} spc_buffer_view;

/* optixPathTracer.h:52-66 (LightTraceParams, 40 B) */
typedef struct spc_light_trace_params {
    int32_t     num_core;
    int32_t     core_padding;
    int32_t     M;
    int32_t     M_per_core;
    spc_vertex* ans;          /* device: BDPTVertex[num_core*core_padding]                     */
    uint8_t*    validState;   /* device: bool[num_core*core_padding]                           */
    int32_t     launch_frame;
    int32_t     _pad;
} spc_light_trace_params;

/* optixPathTracer.h:76-88 (PreTraceParams, 32 B) */
typedef struct spc_pretrace_params {
    int32_t num_core;
    int32_t padding;
    int32_t iteration;
    int32_t _pad;
    void*   paths;            /* device: TrainData::pathInfo_sample[num_core]          (48 B)  */
    void*   conns;            /* device: TrainData::pathInfo_node[num_core*padding]    (92 B)  */
} spc_pretrace_params;

/* optixPathTracer.h:372-383 (TrainData::pathInfo_sample = preTracePath, 48 B): one NEE training path */
typedef struct spc_train_path {
    spc_float3 contri;
    float      sample_pdf;
    float      fix_pdf;
    int32_t    begin_ind;     /* [begin_ind, end_ind) into the connection array */
    int32_t    end_ind;
    int32_t    choice_id;
    int32_t    pixel_x, pixel_y;   /* int2 pixel_id (8-aligned) */
    uint8_t    valid;
    uint8_t    _pad[7];
} spc_train_path;

/* optixPathTracer.h:325-371 (TrainData::pathInfo_node = preTraceConnection, 92 B): one split of a training path
 * into an eye prefix ending at A and a light suffix starting at B */
typedef struct spc_train_conn {
    spc_float3 A_position, B_position, A_dir, B_dir, A_normal, B_normal;
    float      peak_pdf;      /* pdf(eye prefix) * contribution(light suffix) */
    int32_t    path_id;
    int32_t    label_A;       /* eye depth until node_label() writes the eye subspace id */
    int32_t    label_B;
    uint8_t    valid;
    uint8_t    light_source;
    uint8_t    _pad[2];
} spc_train_conn;

/* optixPathTracer.h:89-97 (SubspaceSampler, 40 B) */
typedef struct spc_subspace_sampler {
    const spc_vertex* LVC;
    spc_subspace*     subspace;
    float*            cmfs;
    int32_t*          jump_buffer;
    int32_t           vertex_count;
    int32_t           path_count;
} spc_subspace_sampler;

/* optixPathTracer.h:166-190 (subspaceMacroInfo, 40 B).  `subspaceNum` is unused by the reference
 * (NUM_SUBSPACE is a #define, optixPathTracer.h:31); here it carries the runtime K (0 -> ctx K). */
typedef struct spc_subspace_macro_info {
    int32_t        subspaceNum;
    int32_t        _pad;
    spc_tree_node* eye_tree;
    spc_tree_node* light_tree;
    float*         Q;
    float*         CMFGamma;
} spc_subspace_macro_info;

/* optixPathTracer.h:98-137 (envInfo, 56 B) -- environment lighting is unfinished in the reference
 * (readme.md:28) and out of scope; the block is carried for layout only and `valid` must be 0. */
typedef struct spc_env_info {
    uint64_t   tex;
    float*     cmf;
    float      r;
    spc_float3 center;
    int32_t    size, width, height, divLevel, ssBase;
    uint8_t    valid;
    uint8_t    _pad[3];
} spc_env_info;

/* optixPathTracer.h:191-199 (PTParams = MyParams : whitted::LaunchParams, whitted.h:64-84), 352 B */
typedef struct spc_params {
    uint32_t                width;
    uint32_t                height;
    uint32_t                subframe_index;
    uint32_t                _pad0;
    spc_float4*             accum_buffer;   /* device float4[W*H], caller-allocated               */
    uint32_t*               frame_buffer;   /* device uchar4[W*H], caller-allocated (may be NULL) */
    int32_t                 max_depth;      /* 0 -> the reference's literal 50 (raygen.cu:361)    */
    spc_float3              eye, U, V, W;
    uint32_t                _pad1;
    spc_buffer_view         lights;         /* ignored: lights come from spc_scene_upload         */
    spc_buffer_view         materials;      /* ignored: materials come from spc_scene_upload      */
    spc_float3              miss_color;
    uint32_t                _pad2;
    uint64_t                handle;         /* OptixTraversableHandle: ignored                    */
    spc_light_trace_params  lt;
    spc_subspace_sampler    sampler;
    spc_pretrace_params     pre_tracer;
    spc_subspace_macro_info subspace_info;
    spc_env_info            sky;
} spc_params;

/* ------------------------------------------------------------------------------------------ */
/* scene ingest (replaces sutil::Scene::finalize -> buildMeshAccels/buildInstanceAccel,        */
/* sutil/Scene.cpp:731,943,1260, and the HostToDeviceBuffer uploads of scene_shift.cpp:187-328)*/
/* ------------------------------------------------------------------------------------------ */
typedef struct spc_mesh {
    const float*    positions;    /* float3[n_vertices], stride 12 B (scene_shift.cpp:214)        */
    const uint32_t* indices;      /* uint32[3*n_triangles], stride 12 B per triangle (:217)       */
    const float*    texcoords;    /* float2[n_vertices] or NULL (reference zero-fills, :203-206)  */
    uint32_t        n_vertices;
    uint32_t        n_triangles;
    int32_t         material_id;  /* index into materials[]; ignored when light_id >= 0           */
    int32_t         light_id;     /* >= 0: this mesh is the 2-triangle quad of lights[light_id]
                                     (single sided, back-face culled for closest-hit rays,
                                     sutil/Scene.cpp:1030,1085 + cuProg.h:402); -1 otherwise       */
} spc_mesh;

typedef struct spc_texture {
    const uint8_t* rgba;          /* RGBA8, row-major, as decoded by stbi_load(...,STBI_rgb_alpha) */
    int32_t        width, height;
} spc_texture;

/* one ray of a wavefront batch: 32 B in */
typedef struct spc_ray {
    float ox, oy, oz, tmin;
    float dx, dy, dz, tmax;
} spc_ray;

/* closest-hit record: 16 B out.  prim = global triangle index in upload order (mesh order, then
 * triangle order inside the mesh), -1 on a miss (then t,u,v are 0). */
typedef struct spc_hit {
    float   t, u, v;
    int32_t prim;
} spc_hit;

typedef struct spc_bvh_stats {
    uint32_t n_triangles;
    uint32_t n_nodes;           /* 8-wide compressed nodes (80 B each)                            */
    uint32_t n_bvh2_nodes;
    uint32_t max_depth;         /* of the 8-wide tree                                              */
    float    sah_cost;          /* SAH cost of the 8-wide tree (c_node=1, c_tri=0.3)               */
    float    build_ms;          /* device time of the whole build                                  */
    uint64_t bytes_nodes;
    uint64_t bytes_triangles;
} spc_bvh_stats;

/* traversal work counters for the roofline's algorithmic bytes */
typedef struct spc_trace_counters {
    uint64_t rays;
    uint64_t nodes_visited;     /* 8-wide nodes fetched (80 B each)                                */
    uint64_t tris_tested;       /* triangles fetched+tested (48 B each)                            */
} spc_trace_counters;

typedef struct spc_context spc_context;

enum { SPC_RAYFLAG_NONE = 0, SPC_RAYFLAG_CULL_BACK_FACING = 1 };

/* -------------------------------- lifetime ------------------------------------------------- */
/* Replaces Scene::createContext (sutil/Scene.cpp:856-880).  K = number of subspaces
 * (NUM_SUBSPACE, optixPathTracer.h:31), K_light = emitter subspaces (NUM_SUBSPACE_LIGHTSOURCE,
 * :32), connections = CONNECTION_N (:37).  Pass 0 for the reference defaults (1000, int(0.2*K), 3). */
SPC_API int  spc_create(int device, int K, int K_light, int connections, spc_context** out);
SPC_API void spc_destroy(spc_context* ctx);
SPC_API const char* spc_last_error(void);
SPC_API const char* spc_version(void);
/* All work of `ctx` is enqueued on `cuda_stream` (a cudaStream_t; NULL = the legacy default
 * stream the reference uses everywhere, optixPathTracer.cpp:506). */
SPC_API int  spc_set_stream(spc_context* ctx, void* cuda_stream);
SPC_API int  spc_synchronize(spc_context* ctx);   /* CUDA_SYNC_CHECK, sutil/Exception.h:115 */

/* -------------------------------- scene + BVH ---------------------------------------------- */
/* Host pointers in, device copies + BVH out.  Triangle order defines the prim ids used everywhere. */
SPC_API int  spc_scene_upload(spc_context* ctx,
                              const spc_mesh* meshes, int n_meshes,
                              const spc_pbr* materials, int n_materials,
                              const spc_light* lights, int n_lights,
                              const spc_texture* textures, int n_textures);
/* A second context on the same device uses `owner`'s uploaded scene (geometry, materials, lights, textures, BVH: all read-only after the
 * upload) instead of holding a copy: frame lanes (host/spcbpt_main.cpp --lanes) need one replica per GPU, not one per lane.  No counterpart
 * in the reference (one context, one scene).  `owner` must outlive `ctx` and keep its scene while `ctx` renders. */
SPC_API int  spc_scene_share(spc_context* ctx, spc_context* owner);
SPC_API int  spc_bvh_stats_get(spc_context* ctx, spc_bvh_stats* out);

/* -------------------------------- wavefront ray batches ------------------------------------ */
/* Replaces optixTrace for closest-hit rays (cuProg.h:384-461): nearest hit with tmin < t < tmax,
 * ties -> lowest prim id; ray_flags SPC_RAYFLAG_CULL_BACK_FACING culls back faces of single-sided
 * (emitter) triangles only, like OPTIX_RAY_FLAG_CULL_BACK_FACING_TRIANGLES with the reference's
 * per-instance DISABLE_TRIANGLE_FACE_CULLING.  Host-buffer variant copies H2D/D2H inside. */
SPC_API int  spc_trace_batch(spc_context* ctx, const spc_ray* rays_host, int64_t n, int ray_flags, spc_hit* hits_host);
SPC_API int  spc_trace_batch_device(spc_context* ctx, const spc_ray* rays_dev, int64_t n, int ray_flags, spc_hit* hits_dev);
/* Replaces visibilityTest's optixTrace (cuProg.h:463-487): out[i] = 1 when NO triangle is hit on
 * (tmin, tmax) (i.e. "visible"), else 0.  No face culling (OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT). */
SPC_API int  spc_occlusion_batch(spc_context* ctx, const spc_ray* rays_host, int64_t n, uint8_t* visible_host);
SPC_API int  spc_occlusion_batch_device(spc_context* ctx, const spc_ray* rays_dev, int64_t n, uint8_t* visible_dev);
/* Instrumented (slower) copies of the two kernels above that also count nodes/triangles fetched;
 * results are identical.  Used only to state the algorithmic bytes per ray. */
SPC_API int  spc_trace_batch_counted(spc_context* ctx, const spc_ray* rays_dev, int64_t n, int ray_flags, spc_hit* hits_dev, spc_trace_counters* out_host);
SPC_API int  spc_occlusion_batch_counted(spc_context* ctx, const spc_ray* rays_dev, int64_t n, uint8_t* visible_dev, spc_trace_counters* out_host);
/* -------------------------------- ray-batch producers ---------------------------------------- */
/* Pinhole primaries of one subframe (the raygen part of __raygen__SPCBPT / __raygen__pinhole,
 * raygen.cu:321-344): seed = tea<4>(y*W+x, subframe), jitter (0.5,0.5) at subframe 0 else two rnd
 * draws, dir = normalize(d.x*U + d.y*V + W), tmin 1e-3 (SCENE_EPSILON, cuProg.h:39), tmax 1e16.
 * cam12 = eye,U,V,W (MyParams fields, whitted.h:75-78).  rays_dev: device spc_ray[width*height]. */
SPC_API int  spc_gen_camera_rays(spc_context* ctx, const float* cam12, int width, int height, int subframe, spc_ray* rays_dev);
/* Traversal-microbench ray sets derived from a primary batch and its hits (BASELINE.md section 3,
 * config 2): kind 1 = cosine-hemisphere bounce rays (seed tea<4>(i,1)), kind 2 = shadow rays to a
 * uniform point of light 0 (seed tea<4>(i,2), interval [1e-3, len-1e-3] as cuProg.h:466-475). */
SPC_API int  spc_gen_bench_rays(spc_context* ctx, int kind, const spc_ray* rays_in_dev, const spc_hit* hits_in_dev, int64_t n,
                                spc_ray* rays_out_dev, void* reserved);
/* -------------------------------- launch seam ----------------------------------------------- */
/* The reference copies its MyParams to the device before every launch (optixPathTracer.cpp:495-500,
 * 527-532,616-621) and reads it from the `params` module symbol (cuProg.h:65-67).  spc_set_params takes
 * the same 352-byte struct by value; every pointer inside is a DEVICE pointer owned by the caller
 * (accum_buffer, frame_buffer, lt.ans, lt.validState, pre_tracer.paths/conns, subspace_info.*, sampler.*).
 * `lights`, `materials` and `handle` are ignored (they come from spc_scene_upload). */
SPC_API int  spc_set_params(spc_context* ctx, const spc_params* params);
/* Replaces Scene::switchRaygen(name) + optixLaunch(pipeline, 0, d_params, sizeof(MyParams), sbt, w, h, 1)
 * (sutil/Scene.cpp:1642-1789; optixPathTracer.cpp:502-512, 534-544, 612-632).  Launch sizes as in the
 * reference: SPCBPT_EYE / PT (width, height) = image size; LIGHT_TRACE (lt.num_core, 1); PRETRACE
 * (pre_tracer.num_core, 1).  Asynchronous on the context's stream. */
enum { SPC_LAUNCH_PT = 0, SPC_LAUNCH_SPCBPT_EYE = 1, SPC_LAUNCH_LIGHT_TRACE = 2, SPC_LAUNCH_PRETRACE = 3 };
SPC_API int  spc_launch(spc_context* ctx, int kind, int width, int height);
/* by-name form of the same call: "pt" | "SPCBPT_eye" | "light trace" | "pretrace" (optixPathTracer.cpp:88,502,534,612) */
SPC_API int  spc_launch_named(spc_context* ctx, const char* raygen_name, int width, int height);
/* Multi-GPU sample partition: the eye / pt seeds become tea<4>(pixel, subframe_index + offset) so that ranks rendering the
 * same local subframe indices draw different samples; each rank accumulates its own running mean and the means are
 * reduced over NCCL at read-out.  offset 0 (default) = the reference's streams. */
SPC_API int  spc_set_seed_offset(spc_context* ctx, uint32_t offset);
/* General form: seed = tea<4>(pixel, subframe_index * stride + offset); (0, 1) = the reference.  With (lane, n_lanes) a context that
 * numbers ITS subframes 0,1,2,... (so that its running mean weights stay 1/(n+1), raygen.cu:430-437) draws exactly the samples of the
 * global subframes lane, lane + n_lanes, ...: this is how several contexts on one GPU render alternate subframes concurrently
 * (host/spcbpt_main.cpp --lanes) and how ranks partition subframes across GPUs. */
SPC_API int  spc_set_seed_mapping(spc_context* ctx, uint32_t offset, uint32_t stride);
/* Tile partition of the image (the other way to shard a frame; sutil/WorkDistribution.h:34-91 StaticWorkDistribution, which the reference
 * ships but never wires up): with num_gpus > 1 the SPCBPT_eye launch renders only the pixels of this GPU's 8 x 4 tiles (strips of num_gpus
 * tiles, the GPU's tile shifted by one per strip row) and leaves the other pixels of accum_buffer untouched (the caller zero-fills it).
 * A pixel's estimate depends on its own path and on the light-vertex cache only, so ranks that trace the same light paths (same
 * lt.launch_frame) produce, tile by tile, exactly the single-GPU image: read-out is spc_reduce_accum with weight 1.  Latency, not
 * throughput, scales (strong scaling: every rank still traces the light paths).  (0, 1) = the whole image (default). */
SPC_API int  spc_set_tile_partition(spc_context* ctx, int gpu_idx, int num_gpus);
/* Resident thread blocks per SM of the persistent traversal kernels launched by this context (0 = default: as many as fit, 9).
 * No counterpart in the reference (OptiX schedules its own launches).  A context that shares the GPU with other contexts (frame
 * lanes) leaves room for their small latency-bound kernels by asking for fewer: 7 measured best on the shipped scene with 4 lanes
 * (profiles/r1e_summary.md).  Results do not depend on it. */
SPC_API int  spc_set_trace_blocks(spc_context* ctx, int blocks_per_sm);
/* Read-out of a sample-partitioned render: out = sum_k weights[k] * accum_dev[k] (device float4[n_pixels] each, summed in the order
 * given) and, when out_frame_dev is not NULL, the tone-mapped sRGB uchar4 image of it (ToneMap + make_color, raygen.cu:50-58,
 * cuda/helpers.h:35-67 -- what the eye pass writes to MyParams::frame_buffer).  accum_dev / weights are HOST arrays of n entries. */
SPC_API int  spc_merge_accum(spc_context* ctx, const spc_float4* const* accum_dev, const float* weights, int n, int n_pixels,
                             spc_float4* out_accum_dev, uint32_t* out_frame_dev);
/* Named integer switches of a context (no counterpart in the reference: its equivalents are compile-time #defines in
 * optixPathTracer.h:31-41).  Unknown names fail with SPC_ERR_INVALID.  Results are bit-identical under every switch unless stated.
 *   "reference_search"  1: the eye pass uses the reference's own bisect (binary_sample, cuProg.h:245-264) and walks the
 *                          reference-layout trees instead of guide tables / compact trees / cached cross labels (test switch)
 *   "blocking_sync"     1: the eye pass's lagged queue-size read-backs block the host thread instead of spinning
 *   "count_canonical"   1: the *_counted entry points count the nodes / triangles of the strict front-to-back, t-pruned traversal
 *                          (one ray per lane, one node per step: the traversal SURVEY.md section 8d defines the algorithmic bytes
 *                          by) instead of the visits of the production kernel's own schedule
 *   "light_trace_mode"  0 (default): the reference's RNG streams -- the M_per_core paths of a core share two streams
 *                          (raygen.cu:624-628), so a core is traced serially and the LVC equals the reference's bit for bit;
 *                       1: one lane per light path with per-path streams (seeds tea<4>(0x80000000 | path, launch_frame) for the
 *                          raygen side, tea<4>(0x40000000 | path, launch_frame) for the hit side), vertices
 *                          packed densely in path order: same estimator and distribution, different random numbers (NOT bit-comparable
 *                          with mode 0), an order of magnitude faster
 *   "tail_threshold"    live-path count below which the eye pass stops launching per-bounce wavefront stages and finishes every
 *                          surviving path in one kernel (0 = default 131072, -1 = never); frames are bit-identical for every value
 *   "sort_hits"         1: from the second bounce on, the wavefront queue is re-ordered by the Morton code of the hit points after the
 *                          closest-hit pass (counting sort), so that shading, subspace sampling, shadow rays and the next bounce's rays
 *                          run on spatially coherent entries; frames are bit-identical with and without
 *   "stage_timing"      1: the eye pass brackets every stage of every bounce with CUDA events (slower: for spc_eye_stats_get) */
SPC_API int  spc_set_option(spc_context* ctx, const char* name, int64_t value);
SPC_API int  spc_get_option(spc_context* ctx, const char* name, int64_t* value);
/* Work counters of the LAST eye pass of this context (no counterpart in the reference; bench.py's in-frame Mrays/s).  The call
 * synchronises the context's stream.  stage_ms is filled (timed = 1) when the pass ran under option "stage_timing". */
enum { SPC_STAGE_TRACE = 0, SPC_STAGE_SHADE, SPC_STAGE_SAMPLE, SPC_STAGE_SHADOW, SPC_STAGE_CONNECT, SPC_STAGE_GATHER, SPC_STAGE_OTHER, SPC_STAGE_TOTAL };
typedef struct spc_eye_stats {
    int32_t  bounces;              /* bounces launched                                                      */
    int32_t  timed;
    uint64_t closest_rays;         /* closest-hit rays traced (sum over bounces of the live paths)          */
    uint64_t shadow_slots;         /* connection slots the occlusion kernel scanned (closest_rays * C)      */
    uint64_t shadow_rays;          /* slots that carried a shadow ray (a light vertex was sampled)          */
    uint64_t visible_connections;  /* shadow rays that found no occluder = connections evaluated            */
    float    stage_ms[8];          /* indexed by SPC_STAGE_*: device time per stage summed over the bounces */
} spc_eye_stats;
SPC_API int  spc_eye_stats_get(spc_context* ctx, spc_eye_stats* out);
/* optional parity dumps of the eye pass: per pixel, the primitive id of the primary hit (-1 miss) and the
 * subspace id of the first eye vertex (-1 none).  Device int[W*H] each, or NULL to disable. */
SPC_API int  spc_set_debug_outputs(spc_context* ctx, int32_t* first_prim_dev, int32_t* first_label_dev);

/* -------------------------------- post-processing seam (MyThrustOp) -------------------------- */
/* MyThrustOp::LVC_Process(vertices, validState, countRange) (cuda_thrust/device_thrust.cu:241-332): bins the
 * valid light vertices by subspace id in slot order and builds the per-subspace cmf tables, entirely on the
 * device.  The arrays behind out->subspace / cmfs / jump_buffer are owned by the context and stay valid
 * until the next call (the reference's function-static device_vectors behave the same way). */
SPC_API int  spc_lvc_process(spc_context* ctx, const spc_vertex* lvc_dev, const uint8_t* valid_dev, int count_range,
                             spc_subspace_sampler* out_host);

/* ---- subspace training (the rest of the MyThrustOp seam, cuda_thrust/device_thrust.h:109-135).  Like the reference's
 * library these calls are stateful: the context owns the accumulated training set (neat_paths / neat_conns), Q, Gamma
 * and the trainer's arrays; returned device pointers stay valid until the next call of the same function. ---- */
/* valid_sample_gather(raw_paths, maxPathSize, raw_conns, maxConns) (:457-493): appends the valid paths of one pretrace
 * launch (and their connections) to the training set; *sample_count = number appended. */
SPC_API int  spc_valid_sample_gather(spc_context* ctx, const spc_train_path* raw_paths_dev, int max_paths,
                                     const spc_train_conn* raw_conns_dev, int max_conns, int* sample_count);
SPC_API int  spc_sample_reweight(spc_context* ctx);                                                 /* sample_reweight() (:574-623) */
/* get_weighted_point_for_tree_building(eye_side, max_size) (:494-527) -> host array; *n = required count */
SPC_API int  spc_get_tree_points(spc_context* ctx, int eye_side, int max_size, spc_divide_weight* out_host, int cap, int* n);
/* eye_tree_to_device / light_tree_to_device (:539-552): uploads a host tree, returns the device copy owned by the context */
SPC_API int  spc_tree_to_device(spc_context* ctx, int eye_side, const spc_tree_node* nodes_host, int n, spc_tree_node** dev_out);
/* preprocess_getQ(vertices, validState, countRange, Q) (:347-409): folds one light-trace launch into the running Q
 * estimate (reset != 0 restarts it, like passing a null Q); *acc_paths = accumulated light-path count */
SPC_API int  spc_preprocess_getQ(spc_context* ctx, const spc_vertex* lvc_dev, const uint8_t* valid_dev, int count_range, int reset,
                                 float** Q_dev, int* acc_paths);
SPC_API int  spc_Q_zero_handle(spc_context* ctx);                                                   /* Q_zero_handle (:335-346) */
SPC_API int  spc_node_label(spc_context* ctx, const spc_tree_node* eye_tree_dev, const spc_tree_node* light_tree_dev);   /* node_label (:569-573) */
SPC_API int  spc_build_optimal_E_train_data(spc_context* ctx, int n_samples);                       /* (:3261-3325) */
SPC_API int  spc_preprocess_getGamma(spc_context* ctx, float** gamma_dev);                          /* (:627-667) */
/* train_optimal_E(E_ptr) (:3327-3344): lr 0.01, batch 20000, 1 epoch when the arguments are 0 */
SPC_API int  spc_train_optimal_E(spc_context* ctx, int batch_size, int epochs, float lr, float** gamma_dev,
                                 float* loss_per_batch_host, int loss_cap, int* n_batches);
SPC_API int  spc_Gamma2CMFGamma(spc_context* ctx, const float* gamma_dev, float** cmf_dev);          /* (:3406-3433) */
/* parity dumps / bookkeeping of the training set */
SPC_API int  spc_train_set_size(spc_context* ctx, int* n_paths, int* n_conns);
SPC_API int  spc_train_set_read(spc_context* ctx, spc_train_path* paths_host, spc_train_conn* conns_host);
SPC_API int  spc_train_data_read(spc_context* ctx, int* N, int* M, float* outlier_threshold, float* f_square, float* pdf0, int* P2N,
                                 float* peak, int* label_E, int* label_P);
SPC_API int  spc_train_reset(spc_context* ctx);
/* plain device<->host copies on the context's device (so that a C host needs no CUDA runtime of its own) */
SPC_API int  spc_device_alloc(spc_context* ctx, size_t bytes, void** dev_out);
SPC_API int  spc_device_free(spc_context* ctx, void* dev);
SPC_API int  spc_upload(spc_context* ctx, void* dev, const void* host, size_t bytes);
SPC_API int  spc_download(spc_context* ctx, void* host, const void* dev, size_t bytes);

/* classTree::buildTreeBaseOnExistSample()(samples, subspaceSize, labelBias) (decisionTree/classTree_host.h:302-431):
 * host-side build of a classification tree from weighted sample points (the reference runs it on the host as well,
 * optixPathTracer.cpp:563-567).  Writes at most `cap` nodes to `out` and returns the node count (nothing is
 * written when it exceeds cap) or a negative spc_status; *max_label receives the largest label used. */
SPC_API int  spc_build_tree(const spc_divide_weight* samples, int n, int K, int label_bias, spc_tree_node* out, int cap, int* max_label);

/* -------------------------------- multi-GPU (NCCL over NVLink / NVSwitch) ------------------------ */
/* One process per GPU, one context per process; every rank holds a replica of scene + BVH and renders its own subframes
 * (spc_set_seed_mapping).  The reference is single-GPU: these calls have no counterpart there (its unused tile partition is
 * sutil/WorkDistribution.h:34-91); SURVEY.md section 8b/8e lists them.  Rank 0 creates the id (ncclGetUniqueId), the host ships the
 * 128 bytes to the other ranks by any means (a file, a pipe, MPI, torch.distributed), every rank calls spc_comm_init.
 * Once a context has a communicator of world > 1 the training calls exchange what the single-GPU code sums over the whole training
 * set (every rank holds a SHARD of the training paths and of the Q light-trace launches):
 *   spc_sample_reweight            all-reduces the 10x10-pixel reweighting grid
 *   spc_allreduce_training_stats   folds the per-rank Q estimates (call once after the last spc_preprocess_getQ)
 *   spc_build_optimal_E_train_data takes rank 0's outlier threshold
 *   spc_preprocess_getGamma        all-reduces the Gamma histogram before its row normalisation
 *   spc_train_optimal_E            all-reduces the K x K gradient of every Adam step (global batch = world x batch_size; every rank
 *                                  must hold the same number of batches)
 * so that all ranks end with the same Q / Gamma, equal (up to fp32 summation order) to a single-GPU run over the union of the shards. */
enum { SPC_COMM_ID_BYTES = 128 };
enum { SPC_COMM_I32 = 0, SPC_COMM_F32 = 1, SPC_COMM_F64 = 2 };
enum { SPC_COMM_SUM = 0, SPC_COMM_MIN = 1, SPC_COMM_MAX = 2 };
SPC_API int  spc_comm_unique_id(void* id_out /* SPC_COMM_ID_BYTES */);
SPC_API int  spc_comm_init(spc_context* ctx, int rank, int world, const void* id);
SPC_API int  spc_comm_destroy(spc_context* ctx);
SPC_API int  spc_comm_info(spc_context* ctx, int* rank, int* world);
SPC_API int  spc_comm_barrier(spc_context* ctx);
/* small HOST arrays (counts, timings) all-reduced in place / broadcast from root, staged through the device */
SPC_API int  spc_comm_allreduce_host(spc_context* ctx, void* host_buf, int count, int dtype, int op);
SPC_API int  spc_comm_bcast_host(spc_context* ctx, void* host_buf, size_t bytes, int root);
SPC_API int  spc_allreduce_training_stats(spc_context* ctx);
/* read-out of a sample-partitioned render: accum <- weight * accum summed over ranks into root's buffer (root < 0: into every rank's) */
SPC_API int  spc_reduce_accum(spc_context* ctx, spc_float4* accum_dev, int n_pixels, float weight, int root);
/* import of a training set / Q (the counterparts of spc_train_set_read and of the Q pointer spc_preprocess_getQ returns): restores
 * the state spc_valid_sample_gather / spc_preprocess_getQ build up, e.g. to re-train from a saved set */
SPC_API int  spc_train_set_write(spc_context* ctx, const spc_train_path* paths_host, int n_paths, const spc_train_conn* conns_host, int n_conns);
SPC_API int  spc_train_Q_write(spc_context* ctx, const float* Q_host, int acc_paths);

/* Device-side tree build (csrc/tree_build.cu): the same trees as spc_build_tree / the reference's host builder, bit for bit, built on
 * the GPU -- nearest-centre labelling (N x K) and a level-synchronous octree split.
 * spc_build_tree_gpu: spc_build_tree with host samples in / host nodes out on the GPU of `ctx` (*n_nodes = node count; nodes are
 *   written when they fit `cap`).
 * spc_build_tree_from_training_set: get_weighted_point_for_tree_building + buildTreeBaseOnExistSample + eye/light_tree_to_device in
 *   one call with the weighted points never leaving the device (optixPathTracer.cpp:563-571); the tree becomes the context's eye /
 *   light tree like after spc_tree_to_device; nodes_host (optional, cap entries) receives a host copy for checkpoints. */
SPC_API int  spc_build_tree_gpu(spc_context* ctx, const spc_divide_weight* samples_host, int n, int K, int label_bias, spc_tree_node* out_host, int cap,
                                int* max_label, int* n_nodes);
SPC_API int  spc_build_tree_from_training_set(spc_context* ctx, int eye_side, int max_size, int subspaces, int label_bias, spc_tree_node** dev_out, int* n_nodes,
                                              spc_tree_node* nodes_host, int cap);

/* number of kernels this context has launched so far (bench.py's gpu_launches) */
SPC_API int64_t spc_launch_count(spc_context* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SPCBPT_B200_H */
