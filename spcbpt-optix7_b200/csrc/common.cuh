// common.cuh -- shared declarations of libspcbpt_b200 (sm_100a only; no CPU fallback anywhere).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/spcbpt_b200.h"

namespace spc {

// ---------------------------------------------------------------------------------------------
// error plumbing: C ABI returns codes, message kept per thread (spc_last_error)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
struct CudaFailure { int code; };

#define SPC_CUDA(call)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            spc::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                  \
                           cudaGetErrorString(e__));                                             \
            throw spc::CudaFailure{SPC_ERR_CUDA};                                                \
        }                                                                                        \
    } while (0)

#define SPC_REQUIRE(cond, code, ...)                                                             \
    do {                                                                                         \
        if (!(cond)) {                                                                           \
            spc::set_error(__VA_ARGS__);                                                         \
            throw spc::CudaFailure{code};                                                        \
        }                                                                                        \
    } while (0)

// NVTX range (header-only NVTX3: a no-op unless a profiler is attached): one per API-level stage and per wavefront stage of the
// eye pass, so that an Nsight timeline reads "light trace / LVC_Process / eye pass: bounce b: trace, shade, ..."
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// RAII device buffer (cudaMalloc; buffers are sized for B200's 180 GB, no pooling needed)
template <typename T>
struct DevBuf {
    T*     p = nullptr;
    size_t n = 0;
    bool   owned = true;   // false: a view of another context's buffer (spc_scene_share), never freed or reused for new contents here
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
        n = 0;
        owned = true;
    }
    void alloc(size_t count) {
        if (count <= n && p && owned) return;
        release();
        if (count == 0) count = 1;
        SPC_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void borrow(const DevBuf& o) {
        release();
        p = o.p;
        n = o.n;
        owned = false;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// ---------------------------------------------------------------------------------------------
// compressed 8-wide BVH (80-byte nodes) + 48-byte triangles, single flattened level
// ---------------------------------------------------------------------------------------------
// Node = 5 x float4 (all loads are 128-bit):
//   n0 = { p.x, p.y, p.z, bits(ex | ey<<8 | ez<<16 | imask<<24) }     box origin, biased exponents
//   n1 = { bits(child_base), bits(tri_base), bits(tword), 0 }
//   n2 = { qlox[0..3], qlox[4..7], qloy[0..3], qloy[4..7] }           8-bit quantised child boxes
//   n3 = { qloz[0..3], qloz[4..7], qhix[0..3], qhix[4..7] }
//   n4 = { qhiy[0..3], qhiy[4..7], qhiz[0..3], qhiz[4..7] }
// tword: bits 3s..3s+cnt-1 set for a leaf with cnt (<= 3) triangles in slot s; the triangles of a node are stored in slot order, so
// a triangle's index is tri_base + its rank among the set bits of tword.  Inner children: imask bit s; child index = child_base +
// rank of s in imask.  Empty slots carry an inverted box (qlo 255, qhi 0).
// Triangle = 3 x float4: { v0.xyz, bits(prim) }, { e1.xyz, bits(flags) }, { e2.xyz, 0 } with
// e1 = v1-v0, e2 = v2-v0 (one IEEE subtraction each, as the intersection contract prescribes).
struct Bvh8 {
    DevBuf<float4> nodes;      // 5 per node
    DevBuf<float4> tris;       // 3 per triangle, BVH order
    uint32_t       n_nodes = 0;
    uint32_t       n_tris  = 0;
};

enum : uint32_t { TRI_FLAG_SINGLE_SIDED = 1u };

// de-indexed shading geometry in global prim order (gathered once per hit)
struct SceneGeom {
    DevBuf<float4>   tri_pos;     // 3 per prim: {P0, bits(material)}, {P1, bits(light_id)}, {P2, bits(mesh)}
    DevBuf<float2>   tri_uv;      // 3 per prim
    DevBuf<spc_pbr>  materials;
    DevBuf<float>    mat_log_cc;  // per material: log of the squared clearcoat alpha (see shade.cuh GTR1)
    DevBuf<spc_light> lights;
    DevBuf<uint8_t>  tex_data;    // all textures, RGBA8, concatenated
    DevBuf<int4>     tex_desc;    // {offset_bytes, width, height, 0} per texture
    uint32_t         n_prims = 0;
    int              n_materials = 0, n_lights = 0, n_textures = 0;
    float            scene_lo[3], scene_hi[3];
};

// per-pixel path state + wavefront queues of the eye pass (render.cu), sized on first use
struct EyeBuffers {
    DevBuf<spc_vertex> ev;        // current eye vertex of every path (indexed by pixel)
    DevBuf<float4>     pre;       // {pre-loaded BSDF value toward the sampled direction, pre-loaded singlePdf}
    DevBuf<float4>     res;       // {radiance estimate of this subframe, seed bits}
    DevBuf<spc_ray>    rays[2];   // ping-pong ray queues (queue order)
    DevBuf<int>        queue[2];  // ping-pong pixel ids (queue order)
    DevBuf<int>        queue_ident;   // the first bounce's queue: pixel i at entry i (filled once)
    size_t             ident_pixels = 0;
    DevBuf<spc_hit>    hits;
    DevBuf<spc_ray>    rays_sorted;   // option "sort_hits": the queue of a bounce re-ordered by hit-point Morton code (render.cu)
    DevBuf<int>        queue_sorted;
    DevBuf<spc_hit>    hits_sorted;
    DevBuf<uint32_t>   sort_keys;
    DevBuf<int>        sort_hist;
    DevBuf<spc_ray>    shadow;    // `connections` shadow rays per queue entry
    DevBuf<uint8_t>    visible;
    DevBuf<int>        conn_lvc;  // LVC index of every connection (-1: none)
    DevBuf<float>      conn_pmf;  // path_count * pmf_1 * pmf_2
    DevBuf<float4>     contrib;   // per connection: contribution / pmf / CONNECTION_N (0 when rejected)
    DevBuf<int>        counts;    // counts[b] = live paths entering bounce b
    DevBuf<short>      xlab;      // per pixel: light-tree label of the current eye vertex
    DevBuf<short>      lvc_xlabel;   // per LVC slot: eye-tree label (valid slots only)
    DevBuf<unsigned long long> stat;   // work counters of the last pass (spc_eye_stats_get): shadow rays, visible connections, ...
    int                last_bounces = 0;            // bounces launched by the last pass
    std::vector<cudaEvent_t> stage_ev;              // option "stage_timing": 7 events per bounce + 2 around the pass
    bool               last_timed = false;
    size_t             pixels = 0;
    int                conns = 0;
};

// outputs of the LVC binning (lvc.cu) = MyThrustOp::LVC_Process's SubspaceSampler arrays, owned by the context
struct LvcBuffers {
    DevBuf<spc_subspace> subspace;
    DevBuf<float>        cmfs;
    DevBuf<int>          jump;
    DevBuf<float>        weight;      // per LVC slot
    DevBuf<int>          key;         // per LVC slot: subspace id or -1
    DevBuf<float>        wsorted;
    DevBuf<int>          hist;        // [chunks][K]
    DevBuf<int>          totals;      // [K] + counters
    DevBuf<int>          guide;       // guide tables of the cmfs (shade.cuh): table of subspace b at jump_bias + b, size + 1 entries
    bool                 guide_valid = false;
    int                  n = 0;
};

// the accumulated training set and everything derived from it (train.cu): the file-static state of
// cuda_thrust/device_thrust.cu (neat_paths/neat_conns :428-429, Q_vec :333-334, Gamma_vec :624-625, E_td :3109)
struct TrainBuffers {
    DevBuf<spc_train_path> paths;
    DevBuf<spc_train_conn> conns;
    size_t n_paths = 0, n_conns = 0;
    DevBuf<int> flag_p, flag_c, pos_p, pos_c, scan_sums, scan_sums2, totals;
    DevBuf<spc_divide_weight> tree_pts;
    DevBuf<spc_tree_node> eye_tree, light_tree;
    DevBuf<float4>        eye_ctree, light_ctree;   // compact copies for the device-side walks (shade.cuh "compact trees")
    DevBuf<float> Q;
    int   acc_valid_path = 0;
    bool  has_Q = false;
    int   N = 0, M = 0;
    float outlier_threshold = 0.f;
    DevBuf<float> outlier, f_square, pdf0, peak;
    DevBuf<int>   P2N, label_E, label_P;
    std::vector<int> h_P2N;
    DevBuf<float> gamma, cmf, theta, adam_m, adam_v, E, dE, Esum, loss;
    DevBuf<int>   cmf_guide;   // guide tables of the CDF rows, [K][K+1] (train.cu)
    // deterministic scatter-adds (train.cu): elements sorted by target cell (stable), summed in order per cell
    DevBuf<uint32_t> sort_keys, sort_keys2;
    DevBuf<float>    sort_vals, sort_vals2, path_d, path_loss;
    DevBuf<int>      sort_idx, sort_idx2, node_path;
    DevBuf<uint8_t>  sort_tmp;
};

// spc_set_option switches (include/spcbpt_b200.h documents each)
enum Option { OPT_REFERENCE_SEARCH = 0, OPT_BLOCKING_SYNC, OPT_COUNT_CANONICAL, OPT_STAGE_TIMING, OPT_LIGHT_TRACE_MODE, OPT_TAIL_THRESHOLD, OPT_SORT_HITS, OPT_TRAIN_RESERVE, OPT_COUNT };

struct Context {
    int           device = 0;
    int           K = 1000, K_light = 200, connections = 3;
    cudaStream_t  stream = 0;
    int           sm_count = 148;
    bool          has_scene = false;
    SceneGeom     geom;
    Bvh8          bvh;
    spc_bvh_stats bvh_stats = {};
    int64_t       launches = 0;
    // scratch for host-buffer entry points
    DevBuf<spc_ray> scratch_rays;
    DevBuf<spc_hit> scratch_hits;
    DevBuf<uint8_t> scratch_vis;
    DevBuf<unsigned long long> counters;
    cudaStream_t  pipe_streams[3] = {nullptr, nullptr, nullptr};   // host-buffer batch pipeline (api.cu)
    cudaEvent_t   pipe_event = nullptr;
    DevBuf<void*> merge_ptrs;                    // spc_merge_accum argument staging
    DevBuf<float> merge_w;
    DevBuf<unsigned long long> fetch_counters;   // ray-fetch counters of the persistent traversal kernels (one slot per launch)
    unsigned      fetch_slot = 0;
    int           trace_blocks_per_sm = 0;   // spc_set_trace_blocks: 0 = as many as fit
    // render path
    uint32_t      seed_offset = 0;             // see DevFrame::seed_offset / seed_stride
    uint32_t      seed_stride = 1;
    spc_params    params = {};
    bool          has_params = false;
    EyeBuffers    eye;
    LvcBuffers    lvc;
    int*          dbg_first_prim = nullptr;    // optional device outputs of the eye pass (parity dumps)
    int*          dbg_first_label = nullptr;
    int*          h_pinned = nullptr;          // small pinned staging block for counter read-backs
    cudaEvent_t   eye_events[8] = {};          // lagged queue-size read-backs of the eye pass (render.cu)
    DevBuf<spc_vertex> pretrace_scratch;       // per-lane eye-vertex buffers of the training tracer
    DevBuf<int>   lt_counts;                   // parallel light tracer: per-path vertex counts + offsets (render.cu)
    DevBuf<spc_vertex> lt_scratch;             // ... and its per-path vertex windows (max_depth + 3 vertices each)
    TrainBuffers  train;
    LvcBuffers    bins_tmp;                    // ordered-binning scratch of getQ / sample_reweight
    // per-context one-time setup (function attributes and occupancy are per DEVICE, and spc_create accepts any device)
    bool          lvc_attr_set = false;        // lvc.cu: dynamic shared-memory opt-in of the binning kernels
    int           persist_blocks[2] = {0, 0};  // trace.cu: resident blocks per SM of k_trace_persist<ANYHIT> on this device (0 = not queried)
    int64_t       opt[OPT_COUNT] = {};         // spc_set_option switches (api_render.cu), all 0 by default
    // tile partition of the image across GPUs (spc_set_tile_partition; sutil/WorkDistribution.h): this rank's pixel list
    int           tile_rank = 0, tile_world = 1, n_tile_pixels = 0;
    int           tile_key[4] = {0, 0, 0, 0};
    DevBuf<int>   tile_pixels;
    // multi-GPU (comm.cu): NCCL communicator of this rank, or null (world 1)
    void*         comm = nullptr;
    int           comm_rank = 0, comm_world = 1;
    DevBuf<uint8_t> comm_scratch;
};

void build_bvh(Context& ctx, const float4* d_tri_pos /*3 per prim*/, uint32_t n_prims);
void build_material_tables(Context& ctx);

void launch_trace_closest(Context& ctx, const spc_ray* rays, int64_t n, int flags, spc_hit* hits,
                          unsigned long long* counters /*nullable*/);
void launch_trace_occlusion(Context& ctx, const spc_ray* rays, int64_t n, uint8_t* visible,
                            unsigned long long* counters /*nullable*/);

void launch_trace_closest_q(Context& ctx, const spc_ray* rays, const int* n_dev, int mult, int64_t n_max, int flags, spc_hit* hits);
// the pinhole camera of one subframe (raygen.cu:321-344): the first bounce's closest-hit pass generates its rays itself (ray i = pixel i)
struct CamGen {
    float3   eye, U, V, W;
    unsigned width, height;
    uint32_t sample_index;   // subframe_index * seed_stride + seed_offset
    const int* pixels;       // ray i belongs to pixel pixels[i] (tile partition), or null: pixel i
};
void launch_trace_closest_camera(Context& ctx, const CamGen& cam, int64_t n_pixels, int flags, spc_hit* hits);
void launch_trace_occlusion_q(Context& ctx, const spc_ray* rays, const int* n_dev, int mult, int64_t n_max, uint8_t* visible);

void launch_light_trace(Context& ctx);                       // "light trace" raygen
void launch_eye_pass(Context& ctx, int width, int height);   // "SPCBPT_eye" raygen
void eye_stats(Context& ctx, spc_eye_stats* out);
void merge_accum(Context& ctx, const spc_float4* const* bufs_host, const float* weights_host, int n, int n_pix, spc_float4* out, uint32_t* frame);
void launch_pretrace(Context& ctx);                          // "pretrace" raygen
void launch_pt(Context& ctx, int width, int height);         // "pt" raygen
void lvc_process(Context& ctx, const spc_vertex* lvc, const uint8_t* valid, int n, spc_subspace_sampler* out);
void bin_ordered(Context& c, LvcBuffers& b, int n, int K, int* counters);
int* lvc_bin(Context& c, LvcBuffers& b, const spc_vertex* lvc, const uint8_t* valid, int n);

int    train_gather(Context& c, const spc_train_path* raw_paths, int max_paths, const spc_train_conn* raw_conns, int max_conns);
void   train_reweight(Context& c);
int    train_tree_points(Context& c, int eye_side, int max_size, spc_divide_weight* out_host, int cap);
int    train_tree_points_device(Context& c, int eye_side, int max_size);
int    tree_build_device(Context& c, const spc_divide_weight* samples_dev, int n, int K, int label_bias, DevBuf<spc_tree_node>& out, int* max_label_host);
void   tree_install(Context& c, int eye_side, const spc_tree_node* nodes_host, int n, bool upload);
void   train_node_label(Context& c, const spc_tree_node* eye_tree, const spc_tree_node* light_tree);
int    train_get_Q(Context& c, const spc_vertex* lvc, const uint8_t* valid, int n, int reset);
void   train_Q_zero_handle(Context& c);
void   train_build_data(Context& c, int n_samples);
float* train_get_gamma(Context& c);
float* train_optimal_E(Context& c, int batch_size, int epochs, float lr, float* loss_out_host, int loss_cap, int* n_loss);
float* train_gamma_to_cmf(Context& c, const float* gamma_dev);
void   train_allreduce_Q(Context& c);
// comm.cu: no-ops when the context has no communicator (world 1)
void   comm_allreduce_sum(Context& c, float* dev, size_t n);
void   comm_bcast(Context& c, void* dev, size_t bytes, int root);
// compact copy of a classification tree uploaded by spc_tree_to_device, found by the reference-layout device pointer (or null)
const float4* ctree_lookup(const spc_tree_node* tree_dev);
void ctree_register(const spc_tree_node* tree_dev, const float4* ctree_dev, const void* owner);
const int* gamma_guide_lookup(const float* cmf_gamma, int K);   // guide tables of a CDF built by train_gamma_to_cmf, or null
void gamma_guide_forget(const void* owner, bool trees_too);   // context teardown: trees_too = true

}  // namespace spc

// the opaque handle of the C ABI
struct spc_context {
    spc::Context c;
};
