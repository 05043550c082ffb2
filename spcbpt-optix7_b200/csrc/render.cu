// render.cu -- the two per-frame passes of SPCBPT as hand-written CUDA for sm_100a:
//   * light trace  (replaces optixLaunch of __raygen__lightTrace, raygen.cu:620-685, + the closest-hit programs
//                   __closesthit__lightSubpath / __closesthit__lightSource_subpath, hit_program.cu:341-438,239-244)
//   * eye pass     (replaces optixLaunch of __raygen__SPCBPT, raygen.cu:319-443, + __closesthit__eyeSubpath /
//                   __closesthit__eyeSubpath_LightSource, hit_program.cu:246-340,62-147)
// The eye pass is a wavefront: per bounce  trace -> shade+classify+sample connections -> shadow rays ->
// connection eval + MIS -> ordered gather, with queue sizes kept on the device (no host round trip per stage).
#include <algorithm>
#include <cstdlib>
#include "shade.cuh"
#include "traverse.cuh"

namespace spc {

DevFrame make_dev_frame(Context& c) {
    DevFrame fr;
    fr.sc.tri_pos = c.geom.tri_pos.p;
    fr.sc.tri_uv = c.geom.tri_uv.p;
    fr.sc.materials = c.geom.materials.p;
    fr.sc.lights = c.geom.lights.p;
    fr.sc.tex_data = c.geom.tex_data.p;
    fr.sc.tex_desc = c.geom.tex_desc.p;
    fr.sc.nodes = c.bvh.nodes.p;
    fr.sc.tris = c.bvh.tris.p;
    fr.sc.mat_log_cc = c.geom.mat_log_cc.p;
    fr.sc.n_lights = c.geom.n_lights;
    fr.sc.n_materials = c.geom.n_materials;
    fr.p = c.params;
    fr.K = c.K;
    fr.connections = c.connections;
    fr.lvc_xlabel = nullptr;
    fr.gamma_guide = nullptr;
    fr.lvc_guide = nullptr;
    fr.eye_ctree = c.has_params ? ctree_lookup(c.params.subspace_info.eye_tree) : nullptr;
    fr.light_ctree = c.has_params ? ctree_lookup(c.params.subspace_info.light_tree) : nullptr;
    fr.max_depth = c.params.max_depth > 0 ? c.params.max_depth : 50;
    fr.seed_offset = c.seed_offset;
    fr.seed_stride = c.seed_stride;
    return fr;
}

// per-material constants of the BSDF, computed once at scene upload by the same device functions the kernels would call
__global__ void k_material_tables(const spc_pbr* __restrict__ mats, int n, float* __restrict__ log_cc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = lerpf(0.1f, 0.001f, mats[i].clearcoatGloss);
    log_cc[i] = cm_logf(a * a);
}
void build_material_tables(Context& c) {
    const int n = c.geom.n_materials;
    c.geom.mat_log_cc.alloc(n);
    if (n) k_material_tables<<<(n + 127) / 128, 128, 0, c.stream>>>(c.geom.materials.p, n, c.geom.mat_log_cc.p);
    SPC_CUDA(cudaGetLastError());
    c.launches++;
}

// =============================================================================================
// light trace, reference streams: one sequential "core" per launch index.  The reference's two RNG
// streams per core (raygen side: 5 draws per path; hit side: 4 draws per bounce, both started from the same
// state, raygen.cu:624-628) make path j of a core depend on the bounce counts of paths < j, so a core is
// inherently serial; the 1000 cores run one per warp (lane 0) to keep every SM scheduler busy without
// intra-warp divergence.  This kernel is latency-bound by construction and is meant to run on a side
// stream under the previous frame's eye pass.
// =============================================================================================
constexpr int kLtWarps = 4;

// One light sub-path (the body of the `while (true)` of __raygen__lightTrace, raygen.cu:636-679, with the closest-hit programs
// inlined): draws from `seed` (raygen side: light pick, point, direction) and `hit_seed` (hit side: BSDF sample + Russian roulette),
// hands every vertex to emit(vertex) in path order; emit returns false when the output is full, which ends the path.
// Returns false when emit refused a vertex.
template <class Emit>
__device__ __forceinline__ bool light_path(const DevFrame& fr, uint32_t& seed, uint32_t& hit_seed, uint2* stack, int sstride, const TravLut& lut, Emit& emit) {
    unsigned cn = 0, ct = 0;
    const int li = pick_light(fr, seed);
    LightSample ls;
    {
        const float r1 = rnd(seed);
        const float r2 = rnd(seed);
        light_reverse_sample(fr, li, r1, r2, ls);   // lightSample::operator(), cuProg.h:602-621
    }
    light_trace_mode(ls, seed);
    float3 ray_direction = ls.direction;
    float3 ray_origin = ls.position;
    Vtx cur;
    vtx_zero(cur);
    init_vertex_from_light_sample(ls, cur);
    float3 pre_flux = f3(0.f);
    float pre_singlePdf = ls.dir_pdf;   // init_lightSubPath_from_lightSample, raygen.cu:196-213
    if (!emit(cur)) return false;
    bool done = false;
    int depth = 0;
    while (true) {
        TravRay r{ray_origin.x, ray_origin.y, ray_origin.z, ray_direction.x, ray_direction.y, ray_direction.z, SPC_SCENE_EPS, 1e16f};
        TravHit h;
        bool pushed = false;
        if (!traverse_bvh8<false, false>(fr.sc.nodes, fr.sc.tris, r, true, stack, sstride, h, cn, ct, lut)) {
            done = true;                                   // __miss__BDPTVertex, raygen.cu:699-704
        } else {
            const LocalGeom g = hit_geometry(fr.sc, h.prim, h.u, h.v);
            if (g.light >= 0) {
                done = true;                               // __closesthit__lightSource_subpath, hit_program.cu:239-244
            } else {
                Vtx mid;
                SurfaceOut so;
                surface_hit(fr, cur, pre_flux, pre_singlePdf, g, h.t, ray_direction, true, hit_seed, mid, so);
                cur = mid;
                pre_flux = so.next_flux;
                pre_singlePdf = so.next_singlePdf;
                ray_direction = so.dir;
                ray_origin = g.P;
                done = so.done;
                pushed = true;
            }
        }
        if (pushed && !emit(cur)) return false;
        if (done || depth > fr.max_depth) break;
        depth += 1;
    }
    return true;
}

__global__ void __launch_bounds__(kLtWarps * 32) k_light_trace_cores(const DevFrame fr, int lanes) {
    __shared__ uint2 s_stack[kSmStack * kLtWarps * 32];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);   // before any thread leaves
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int core = (blockIdx.x * kLtWarps + warp) * lanes + lane;
    const spc_light_trace_params& lt = fr.p.lt;
    if (lane >= lanes || core >= lt.num_core) return;

    uint32_t seed = tea<4>((uint32_t)core, (uint32_t)lt.launch_frame);
    uint32_t hit_seed = seed;   // payload.seed: a copy taken once (raygen.cu:628)
    const unsigned bias = (unsigned)lt.core_padding * (unsigned)core;
    unsigned n_vert = 0, n_path = 0;
    const unsigned cap = (unsigned)lt.core_padding;
    auto emit = [&](const Vtx& v) {   // pushVertexToLVC (raygen.cu:613-619) + the window check of :656,:676
        vtx_store(lt.ans + bias + n_vert, v);
        lt.validState[bias + n_vert] = 1;
        n_vert++;
        return n_vert < cap;
    };
    while (true) {
        if (!light_path(fr, seed, hit_seed, s_stack + threadIdx.x, kLtWarps * 32, s_lut, emit)) break;
        n_path++;
        if (n_path >= (unsigned)lt.M_per_core) break;
    }
    for (unsigned i = n_vert; i < cap; i++) lt.validState[bias + i] = 0;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Parallel light tracer (spc_set_option "light_trace_mode" 1): one lane per light path instead of one per core.  Every path gets
// its own streams, seeds tea<4>(0x80000000 | path, launch_frame) and tea<4>(0x40000000 | path, launch_frame) (the reference couples the 100 paths of a core through two
// shared streams, which is what forces k_light_trace_cores to run them serially): same estimator, same distribution, different
// random numbers -- so frames are NOT bit-comparable with the reference-stream mode (tests/test_light_trace_modes_gpu.py compares
// them statistically).  The vertices are packed densely in path order: the trace pass writes every path's vertices into its own
// scratch window and counts them, an exclusive scan of the counts places the paths, a compaction kernel moves them from a per-path scratch window (max_depth + 3 vertices) to the LVC -- deterministic and
// bit-reproducible (the assignment of paths to lanes is dynamic, the output depends on the path index only).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int kLtPathBlock = 128;
// Persistent lanes: a lane that finishes its path claims the next path index at once (the output depends on the path index only, so
// any assignment of paths to lanes gives the same LVC), and the loop body is ONE bounce of whatever path the lane holds -- a warp
// never idles behind its longest path (path lengths range from 1 to max_depth + 2 vertices).
__global__ void __launch_bounds__(kLtPathBlock) k_light_trace_paths(const DevFrame fr, int n_paths, int* __restrict__ counts, spc_vertex* __restrict__ scratch, int stride,
                                                                    int* __restrict__ next_path) {
    __shared__ uint2 s_stack[kSmStack * kLtPathBlock];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);
    const spc_light_trace_params& lt = fr.p.lt;
    uint2* stack = s_stack + threadIdx.x;
    unsigned cn = 0, ct = 0;
    int p = -1, n_vert = 0, base = 0, depth = 0;
    uint32_t seed = 0, hit_seed = 0;
    Vtx cur;
    float3 pre_flux = f3(0.f), ray_origin = f3(0.f), ray_direction = f3(0.f);
    float pre_singlePdf = 0.f;
    auto emit = [&](const Vtx& v) {
        if (n_vert >= stride) return false;   // (cannot happen: stride = max_depth + 3 vertices)
        vtx_store(scratch + (size_t)base * stride + n_vert, v);
        n_vert++;
        return true;
    };
    while (true) {
        if (p < 0) {
            p = atomicAdd(next_path, 1);
            if (p >= n_paths) break;
            // the two streams of a path (raygen side / hit side, raygen.cu:624-628) get their own seeds: started from one state, as the
            // reference's first path of a core is, the hit side would replay the raygen side's draws on EVERY path
            seed = tea<4>(0x80000000u | (uint32_t)p, (uint32_t)lt.launch_frame);
            hit_seed = tea<4>(0x40000000u | (uint32_t)p, (uint32_t)lt.launch_frame);
            n_vert = 0;
            depth = 0;
            base = p;
            const int li = pick_light(fr, seed);
            LightSample ls;
            const float r1 = rnd(seed);
            const float r2 = rnd(seed);
            light_reverse_sample(fr, li, r1, r2, ls);
            light_trace_mode(ls, seed);
            ray_direction = ls.direction;
            ray_origin = ls.position;
            vtx_zero(cur);
            init_vertex_from_light_sample(ls, cur);
            pre_flux = f3(0.f);
            pre_singlePdf = ls.dir_pdf;
            emit(cur);
        }
        // one bounce of the path this lane holds (the loop body of light_path)
        const TravRay r{ray_origin.x, ray_origin.y, ray_origin.z, ray_direction.x, ray_direction.y, ray_direction.z, SPC_SCENE_EPS, 1e16f};
        TravHit h;
        bool done = false, pushed = false;
        if (!traverse_bvh8<false, false>(fr.sc.nodes, fr.sc.tris, r, true, stack, kLtPathBlock, h, cn, ct, s_lut)) {
            done = true;
        } else {
            const LocalGeom g = hit_geometry(fr.sc, h.prim, h.u, h.v);
            if (g.light >= 0) {
                done = true;
            } else {
                Vtx mid;
                SurfaceOut so;
                surface_hit(fr, cur, pre_flux, pre_singlePdf, g, h.t, ray_direction, true, hit_seed, mid, so);
                cur = mid;
                pre_flux = so.next_flux;
                pre_singlePdf = so.next_singlePdf;
                ray_direction = so.dir;
                ray_origin = g.P;
                done = so.done;
                pushed = true;
            }
        }
        if (pushed && !emit(cur)) done = true;
        if (done || depth > fr.max_depth) {
            counts[p] = n_vert;
            p = -1;
        } else {
            depth += 1;
        }
    }
}
// exclusive scan of the per-path vertex counts (one block; 10^5 paths): tiles of 1024 counts, coalesced, warp-shuffle scan per tile
// with a running carry (a strided per-thread serial walk of ~100 counts plus a serial pass over 1024 partials took 100 us)
__global__ void __launch_bounds__(1024) k_lt_scan(const int* __restrict__ counts, int n, int* __restrict__ offsets, int* __restrict__ total) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + (int)threadIdx.x;
        const int v = i < n ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_warp[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_warp[lane] = wi - w;   // exclusive prefix of the warp sums
        }
        __syncthreads();
        const int carry = s_carry;
        if (i < n) offsets[i] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}
// dense LVC: path p's vertices go to slots offsets[p].. in path order (one warp per path, 15 x 64-bit words per vertex)
__global__ void k_lt_compact(const spc_vertex* __restrict__ scratch, int stride, const int* __restrict__ counts, const int* __restrict__ offsets, int n_paths, int n_slots,
                             spc_vertex* __restrict__ lvc, uint8_t* __restrict__ valid) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n_paths) return;
    const int cnt = counts[p], base = offsets[p];
    const uint2* src = reinterpret_cast<const uint2*>(scratch + (size_t)p * stride);
    uint2* dst = reinterpret_cast<uint2*>(lvc + base);
    const int keep = max(0, min(cnt, n_slots - base));   // LVC full: the tail of the path order is dropped, deterministically
    for (int w = lane; w < keep * 15; w += 32) dst[w] = src[w];
    for (int k = lane; k < keep; k += 32) valid[base + k] = 1;
}
__global__ void k_lt_clear_tail(uint8_t* __restrict__ valid, const int* __restrict__ total, int n_slots) {
    const int first = min(*total, n_slots);
    for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += gridDim.x * blockDim.x) valid[i] = 0;
}

void launch_light_trace(Context& c) {
    NvtxRange range("spc: light trace");
    SPC_REQUIRE(c.has_params, SPC_ERR_INVALID, "spc_launch: spc_set_params has not been called");
    const spc_light_trace_params& lt = c.params.lt;
    SPC_REQUIRE(lt.num_core > 0 && lt.core_padding > 0 && lt.ans && lt.validState, SPC_ERR_INVALID, "spc_launch(light trace): MyParams::lt is not set up");
    SPC_REQUIRE(c.geom.n_lights > 0, SPC_ERR_NO_SCENE, "spc_launch(light trace): the scene has no lights");
    const DevFrame fr = make_dev_frame(c);
    if (c.opt[OPT_LIGHT_TRACE_MODE] == 1) {
        const int n_paths = lt.num_core * lt.M_per_core, n_slots = lt.num_core * lt.core_padding;
        c.lt_counts.alloc((size_t)2 * n_paths + 4);
        int* counts = c.lt_counts.p;
        int* offsets = counts + n_paths;
        int* total = offsets + n_paths;
        int* next_path = total + 1;
        SPC_CUDA(cudaMemsetAsync(next_path, 0, sizeof(int), c.stream));
        const int stride = fr.max_depth + 3;   // the emitter vertex + one vertex per bounce 0 .. max_depth + 1
        c.lt_scratch.alloc((size_t)n_paths * stride);
        const int grid = std::min((n_paths + kLtPathBlock - 1) / kLtPathBlock, c.sm_count * 2);
        k_light_trace_paths<<<grid, kLtPathBlock, 0, c.stream>>>(fr, n_paths, counts, c.lt_scratch.p, stride, next_path);
        k_lt_scan<<<1, 1024, 0, c.stream>>>(counts, n_paths, offsets, total);
        k_lt_compact<<<(n_paths * 32 + 255) / 256, 256, 0, c.stream>>>(c.lt_scratch.p, stride, counts, offsets, n_paths, n_slots, lt.ans, lt.validState);
        k_lt_clear_tail<<<c.sm_count, 256, 0, c.stream>>>(lt.validState, total, n_slots);
        SPC_CUDA(cudaGetLastError());
        c.launches += 4;
        return;
    }
    static const int lanes = []() {
        const char* e = getenv("SPC_LT_LANES");
        return e ? std::max(1, std::min(32, atoi(e))) : 1;
    }();
    const int per_block = kLtWarps * lanes;
    k_light_trace_cores<<<(lt.num_core + per_block - 1) / per_block, kLtWarps * 32, 0, c.stream>>>(fr, lanes);
    SPC_CUDA(cudaGetLastError());
    c.launches++;
}

// =============================================================================================
// eye pass
// =============================================================================================
struct EyeArgs {
    spc_vertex* ev;
    float4*     pre;
    float4*     res;
    float4*     rays_cur;
    float4*     rays_next;
    int*        queue_cur;
    int*        queue_next;
    const float4* hits;
    float4*     shadow;
    const uint8_t* visible;
    int*        conn_lvc;
    float*      conn_pmf;
    float4*     contrib;
    int*        counts;     // counts[bounce] in, counts[bounce+1] out
    int*        first_prim;
    int*        first_label;
    short*      xlab;       // per pixel: light-tree label of the current eye vertex (cross label, shade.cuh)
    unsigned long long* stat;   // work counters: [0] shadow rays, [1] visible connections (spc_eye_stats_get)
    int         bounce;
};

// init_EyeSubpath (raygen.cu:216-231): the vertex at the camera
__device__ __forceinline__ Vtx camera_vertex(float3 eye, float3 d) {
    Vtx v;
    vtx_zero(v);
    v.position = eye;
    v.flux = f3(1.0f);
    v.pdf = 1.0f;
    v.RMIS_pointer = 0;
    v.normal = d;
    v.isOrigin = 1;
    v.depth = 0;
    v.singlePdf = 1.0f;
    return v;
}
// the path's RNG state after the pixel jitter (raygen.cu:332-336)
__device__ __forceinline__ uint32_t first_seed(const DevFrame& fr, int pix) {
    const uint32_t sample_index = fr.p.subframe_index * fr.seed_stride + fr.seed_offset;
    uint32_t seed = tea<4>((uint32_t)pix, sample_index);
    if (sample_index != 0) {
        rnd(seed);
        rnd(seed);
    }
    return seed;
}

// The raygen part of __raygen__SPCBPT (raygen.cu:321-353) has no kernel of its own: the first closest-hit pass generates the camera
// rays itself (trace.cu, k_trace_persist<.., RAYGEN>) and the first k_eye_shade rebuilds ray, camera vertex and RNG state in
// registers (camera_ray / camera_vertex / first_seed); queue entry i of the first bounce is pixel i (a cached identity array).
// That removes 188 B written and 220 B read per pixel and frame.
__device__ __forceinline__ void camera_ray(const DevFrame& fr, int pix, float3& eye, float3& dir) {
    const unsigned W = fr.p.width, H = fr.p.height;
    const uint32_t sample_index = fr.p.subframe_index * fr.seed_stride + fr.seed_offset;   // = subframe_index in the reference
    uint32_t seed = tea<4>((uint32_t)pix, sample_index);
    float jx = 0.5f, jy = 0.5f;
    if (sample_index != 0) {   // make_float2(rnd(seed), rnd(seed)): nvcc evaluates left to right
        jx = rnd(seed);
        jy = rnd(seed);
    }
    eye = ld3(fr.p.eye);
    dir = camera_dir_exact(ld3(fr.p.U), ld3(fr.p.V), ld3(fr.p.W), (unsigned)pix % W, (unsigned)pix / W, W, H, jx, jy);
}
__global__ void k_eye_begin(const EyeArgs a, int n_pix, int n_first, int n_work, int* __restrict__ ident, int fill_ident) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) a.counts[0] = n_first;   // entries of the first bounce's queue: all pixels, or this rank's tiles
    if (i >= n_work || n_work < n_pix) return;   // n_work = n_pix only when there is per-pixel work
    if (fill_ident) ident[i] = i;
    if (a.first_prim) a.first_prim[i] = -1;
    if (a.first_label) a.first_label[i] = -1;
}

// closest-hit programs + connection sampling of one bounce; one lane per live path
__global__ void __launch_bounds__(128) k_eye_shade(const DevFrame fr, const EyeArgs a) {
    const int n = a.counts[a.bounce];
    const int C = fr.connections;
    for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        bool alive = false;
        float4 nro, nrd;
        int pix = 0;
        if (i < n) {
            pix = a.queue_cur[i];
            const float4 hit = a.hits[i];
            const int prim = __float_as_int(hit.w);
            a.conn_lvc[(size_t)i * C] = -1;   // no surface vertex (yet): k_eye_sample fills the slots either way
            if (a.bounce == 0 && a.first_prim) a.first_prim[pix] = prim;
            if (prim < 0 && a.bounce == 0) a.res[pix] = make_float4(0.f, 0.f, 0.f, __uint_as_float(first_seed(fr, pix)));   // a miss: the pixel's result is 0
            if (prim >= 0) {
                const bool first = a.bounce == 0;
                float3 ray_direction, cam_eye = f3(0.f);
                if (first) {
                    camera_ray(fr, pix, cam_eye, ray_direction);
                } else {
                    const float4 rd4 = a.rays_cur[2 * (size_t)i + 1];
                    ray_direction = f3(rd4.x, rd4.y, rd4.z);
                }
                const Vtx last = first ? camera_vertex(cam_eye, ray_direction) : vtx_load(a.ev + pix);
                const float4 pre = first ? make_float4(0.f, 0.f, 0.f, 1.0f) : a.pre[pix];
                const int last_x = first ? -1 : (int)a.xlab[pix];   // written by k_eye_sample of the previous bounce
                float4 res = first ? make_float4(0.f, 0.f, 0.f, __uint_as_float(first_seed(fr, pix))) : a.res[pix];
                uint32_t seed = __float_as_uint(res.w);
                if (first) a.res[pix] = res;   // (the emitter / surface branches below overwrite it where they add to it)
                const LocalGeom g = hit_geometry(fr.sc, prim, hit.y, hit.z);
                Vtx mid;
                if (g.light >= 0) {
                    // __closesthit__eyeSubpath_LightSource + lightStraghtHit (raygen.cu:305-317, :383-388)
                    if (eye_hits_light(fr, last, f3(pre.x, pre.y, pre.z), pre.w, g, hit.x, ray_direction, mid, last_x)) {
                        if (a.bounce == 0 && a.first_label) a.first_label[pix] = mid.subspaceId;
                        const float3 ans = mid.flux / mid.pdf / mid.RMIS_pointer;
                        if (!invalid3(ans)) {
                            res.x += ans.x; res.y += ans.y; res.z += ans.z;
                            a.res[pix] = res;
                        }
                    }
                } else {
                    SurfaceOut so;
                    surface_hit(fr, last, f3(pre.x, pre.y, pre.z), pre.w, g, hit.x, ray_direction, false, seed, mid, so, last_x, true);
                    vtx_store(a.ev + pix, mid);
                    a.pre[pix] = make_float4(so.next_flux.x, so.next_flux.y, so.next_flux.z, so.next_singlePdf);
                    // the CONNECTION_N probabilistic connections of this vertex are drawn by k_eye_sample (marker -2 in slot 0)
                    a.conn_lvc[(size_t)i * C] = -2;
                    res.w = __uint_as_float(seed);
                    a.res[pix] = res;
                    // loop head of the next iteration (raygen.cu:361): payload.done || payload.depth > 50
                    alive = !so.done && !(a.bounce + 1 > fr.max_depth);
                    nro = make_float4(g.P.x, g.P.y, g.P.z, SPC_SCENE_EPS);
                    nrd = make_float4(so.dir.x, so.dir.y, so.dir.z, 1e16f);
                }
            }
        }
        // warp-vote compaction of the survivors into the next queue
        const unsigned ballot = __ballot_sync(0xffffffffu, alive);
        if (ballot) {
            const int lane = threadIdx.x & 31;
            int slot0 = 0;
            if (lane == (__ffs(ballot) - 1)) slot0 = atomicAdd(a.counts + a.bounce + 1, __popc(ballot));
            slot0 = __shfl_sync(0xffffffffu, slot0, __ffs(ballot) - 1);
            if (alive) {
                const int slot = slot0 + __popc(ballot & ((1u << lane) - 1u));
                a.queue_next[slot] = pix;
                a.rays_next[2 * (size_t)slot] = nro;
                a.rays_next[2 * (size_t)slot + 1] = nrd;
            }
        }
    }
}

// CONNECTION_N probabilistic connections per eye vertex (raygen.cu:390-419): stage 1 picks a light subspace from row
// eye-subspace of the Gamma CDF, stage 2 a vertex of that subspace from its cmf; both with the reference's own bisect
// (binary_sample, cuProg.h:245-264).  Split from k_eye_shade: that kernel needs ~100 registers (4 blocks per SM) and its
// run time was the latency of this chain of ~60 dependent 4-byte probes per lane (ncu: issue slots 22 % busy, long-scoreboard
// stalls dominant, profiles/r1e_summary.md).  Here one lane per path needs few registers (full occupancy), and for C <= 4
// the C bisects of a stage run in lockstep, so their probes overlap.  The draws of connection j are numbers 2j and 2j+1
// of the path's stream only while no earlier connection met an empty subspace (which draws one number, raygen.cu:400-403):
// the lockstep path assumes that and falls back to the serial loop in the rare case it does not hold -- same draws, same
// results as the serial loop in every case.
struct ConnPick {
    int   lv;      // LVC index, -1: no connection
    float pmf;     // path_count * pmf2 * pmf1
};

__device__ __forceinline__ void eye_sample_serial(const DevFrame& fr, int eye_subspace, int C, uint32_t& seed, ConnPick* out) {
    const spc_subspace_sampler& S = fr.p.sampler;
    for (int j = 0; j < C; j++) {
        out[j].lv = -1;
        out[j].pmf = 0.f;
        int light_id = 0;
        float pmf1 = 1;
        if (fr.p.subspace_info.light_tree) {
            const float* row = fr.p.subspace_info.CMFGamma + (size_t)eye_subspace * fr.K;
            light_id = fr.gamma_guide ? guided_sample(row, fr.gamma_guide + (size_t)eye_subspace * (fr.K + 1), fr.K, seed, pmf1)
                                      : binary_sample(row, fr.K, seed, pmf1);
        }
        const spc_subspace sub = S.subspace[light_id];
        if (sub.size != 0) {
            float pmf2;
            const float* cm = S.cmfs + sub.jump_bias;
            const int index = (fr.lvc_guide ? guided_sample(cm, fr.lvc_guide + sub.jump_bias + light_id, sub.size, seed, pmf2)
                                            : binary_sample(cm, sub.size, seed, pmf2)) + sub.jump_bias;
            out[j].lv = S.jump_buffer[index];
            out[j].pmf = (float)S.path_count * pmf2 * pmf1;
        }
    }
}

// CT bisects of equal size over the same table, in lockstep (the trip count of the reference's loop depends on the size only
// through r-l, which differs between lanes of the group by at most the rounding of the halving: each keeps its own l, r)
template <int CT>
__device__ __forceinline__ void bisect_lockstep(const float* const* cmf, const int* size, const float* u, int* l_out, float* pmf_out) {
    int l[CT], r[CT], mid[CT];
#pragma unroll
    for (int j = 0; j < CT; j++) { l[j] = 0; r[j] = size[j]; mid[j] = size[j] / 2 - 1; }
    bool any = true;
    while (any) {
        any = false;
        float v[CT];
#pragma unroll
        for (int j = 0; j < CT; j++) v[j] = (r[j] - l[j] > 1) ? __ldg(cmf[j] + mid[j]) : 0.f;
#pragma unroll
        for (int j = 0; j < CT; j++) {
            if (r[j] - l[j] > 1) {
                if (u[j] < v[j]) r[j] = mid[j] + 1;
                else l[j] = mid[j] + 1;
                mid[j] = (l[j] + r[j]) / 2 - 1;
                any |= (r[j] - l[j] > 1);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CT; j++) {
        const float hi = __ldg(cmf[j] + l[j]);
        const float lo = l[j] == 0 ? 0.f : __ldg(cmf[j] + l[j] - 1);
        pmf_out[j] = l[j] == 0 ? hi : hi - lo;
        l_out[j] = l[j];
    }
}

// CT guided searches in lockstep (shade.cuh, "guide tables"): cell, two adjacent table reads, then a binary search over the bracket
template <int CT>
__device__ __forceinline__ void guided_lockstep(const float* const* cmf, const int* const* G, const int* size, const float* u, int* l_out, float* pmf_out) {
    int lo[CT], hi[CT];
#pragma unroll
    for (int j = 0; j < CT; j++) {
        lo[j] = 0;
        hi[j] = 0;
        if (G[j]) {   // null: a dummy slot (no search, index 0)
            const int cell = guide_cell(u[j], size[j]);
            lo[j] = __ldg(G[j] + cell);
            hi[j] = min(__ldg(G[j] + cell + 1), size[j] - 1);
            lo[j] = min(lo[j], hi[j]);
        }
    }
    bool any = true;
    while (any) {   // first i in [lo, hi) with u < cmf[i], else hi
        any = false;
        float v[CT];
        int mid[CT];
#pragma unroll
        for (int j = 0; j < CT; j++) {
            mid[j] = (lo[j] + hi[j]) >> 1;
            v[j] = lo[j] < hi[j] ? __ldg(cmf[j] + mid[j]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < CT; j++) {
            if (lo[j] < hi[j]) {
                if (u[j] < v[j]) hi[j] = mid[j];
                else lo[j] = mid[j] + 1;
                any |= lo[j] < hi[j];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < CT; j++) {
        const float top = __ldg(cmf[j] + lo[j]);
        const float below = lo[j] == 0 ? 0.f : __ldg(cmf[j] + lo[j] - 1);
        pmf_out[j] = lo[j] == 0 ? top : top - below;
        l_out[j] = lo[j];
    }
}

// All C connections of a vertex without a serial dependency.  Connection j draws its first-stage number at position s_j of the path's
// stream, s_0 = 0, s_{j+1} = s_j + (1 if connection j met an empty subspace, else 2) (raygen.cu:390-408): s_j is one of j..2j.  The first
// stage is therefore run for the 2C-1 candidate positions 0..2C-2 at once (lockstep searches, their probes overlap), the actual
// positions are then resolved in registers, and the second stage runs in lockstep for the connections that have one.  Same draws,
// same searches as the serial loop for every pattern of empty subspaces (in the shipped scene ~10 % of the vertices meet one: a
// speculative version that fell back to the serial loop sent 3 lanes of almost every warp down that path, profiles/r1e_summary.md).
// Requires a light tree (without one the first stage draws nothing: serial loop).
template <int CT>
__device__ __forceinline__ bool eye_sample_lockstep(const DevFrame& fr, int eye_subspace, uint32_t& seed, ConnPick* out) {
    if (!fr.p.subspace_info.light_tree) return false;
    constexpr int NC = 2 * CT - 1;   // candidate first-stage positions
    const spc_subspace_sampler& S = fr.p.sampler;
    uint32_t st[2 * CT + 1];         // st[k] = state after k draws
    float d[2 * CT];
    st[0] = seed;
#pragma unroll
    for (int k = 0; k < 2 * CT; k++) {
        uint32_t t = st[k];
        d[k] = rnd(t) * 1.0f;
        st[k + 1] = t;
    }
    // first stage for every candidate position
    int cand_id[NC];
    float cand_pmf[NC];
    {
        const float* row = fr.p.subspace_info.CMFGamma + (size_t)eye_subspace * fr.K;
        const float* cm[NC];
        int sz[NC];
        float u[NC];
#pragma unroll
        for (int c = 0; c < NC; c++) { cm[c] = row; sz[c] = fr.K; u[c] = d[c]; }
        if (fr.gamma_guide) {
            const int* gt[NC];
#pragma unroll
            for (int c = 0; c < NC; c++) gt[c] = fr.gamma_guide + (size_t)eye_subspace * (fr.K + 1);
            guided_lockstep<NC>(cm, gt, sz, u, cand_id, cand_pmf);
        } else {
            bisect_lockstep<NC>(cm, sz, u, cand_id, cand_pmf);
        }
    }
    int cand_size[NC], cand_bias[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        const spc_subspace sub = S.subspace[cand_id[c]];
        cand_size[c] = sub.size;
        cand_bias[c] = sub.jump_bias;
    }
    // resolve the positions (registers only)
    int light_id[CT], size[CT], bias[CT];
    float pmf1[CT], u2[CT];
    int pos = 0;
#pragma unroll
    for (int j = 0; j < CT; j++) {
        light_id[j] = 0; size[j] = 0; bias[j] = 0; pmf1[j] = 1.f; u2[j] = 0.f;
#pragma unroll
        for (int c = j; c <= 2 * j; c++) {
            if (c == pos) {
                light_id[j] = cand_id[c]; size[j] = cand_size[c]; bias[j] = cand_bias[c]; pmf1[j] = cand_pmf[c];
                u2[j] = d[c + 1];
            }
        }
        pos += size[j] != 0 ? 2 : 1;
    }
    // second stage for the connections that have one (an empty subspace gets a one-entry dummy table and no connection)
    const float* cm[CT];
    int sz[CT], idx[CT];
    float pmf2[CT];
#pragma unroll
    for (int j = 0; j < CT; j++) { cm[j] = S.cmfs + bias[j]; sz[j] = max(size[j], 1); }
    if (fr.lvc_guide) {
        const int* gt[CT];
#pragma unroll
        for (int j = 0; j < CT; j++) gt[j] = size[j] != 0 ? fr.lvc_guide + bias[j] + light_id[j] : nullptr;
        guided_lockstep<CT>(cm, gt, sz, u2, idx, pmf2);
    } else {
        bisect_lockstep<CT>(cm, sz, u2, idx, pmf2);
    }
#pragma unroll
    for (int j = 0; j < CT; j++) {
        out[j].lv = -1;
        out[j].pmf = 0.f;
        if (size[j] != 0) {
            out[j].lv = S.jump_buffer[idx[j] + bias[j]];
            out[j].pmf = (float)S.path_count * pmf2[j] * pmf1[j];
        }
    }
    // seed after the draws actually consumed: pos of them
    uint32_t fin = st[CT];
#pragma unroll
    for (int k = CT; k <= 2 * CT; k++)
        if (k == pos) fin = st[k];
    seed = fin;
    return true;
}

template <int CT>
__global__ void __launch_bounds__(128, 8) k_eye_sample(const DevFrame fr, const EyeArgs a) {
    const int n = a.counts[a.bounce];
    const int C = CT > 0 ? CT : fr.connections;
    unsigned n_shadow = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (a.conn_lvc[(size_t)i * C] != -2) {   // miss, emitter hit or no surface vertex: empty slots (an empty-interval shadow ray each)
            for (int j = 0; j < C; j++) {
                a.conn_lvc[(size_t)i * C + j] = -1;
                a.shadow[2 * ((size_t)i * C + j)] = make_float4(0.f, 0.f, 0.f, 1.0f);
                a.shadow[2 * ((size_t)i * C + j) + 1] = make_float4(0.f, 0.f, 1.f, -1.0f);
            }
            continue;
        }
        const int pix = a.queue_cur[i];
        spc_vertex* ev = a.ev + pix;
        const float3 pos = f3(ev->position.x, ev->position.y, ev->position.z);
        // classification of the new vertex (labelUnit::getLabel in __closesthit__eyeSubpath, hit_program.cu:295) and its cross label
        int eye_subspace, cross;
        const float3 nrm = f3(ev->normal.x, ev->normal.y, ev->normal.z);
        if (fr.eye_ctree && fr.light_ctree) ctree_label2(fr.eye_ctree, fr.light_ctree, pos, nrm, eye_subspace, cross);
        else tree_label2(fr.p.subspace_info.eye_tree, fr.p.subspace_info.light_tree, pos, nrm, eye_subspace, cross);
        ev->subspaceId = (short)eye_subspace;
        a.xlab[pix] = (short)cross;
        if (a.bounce == 0 && a.first_label) a.first_label[pix] = eye_subspace;
        uint32_t seed = __float_as_uint(a.res[pix].w);
        ConnPick pick[CT > 0 ? CT : 16];
        bool done = false;
        if (CT > 0) done = eye_sample_lockstep<(CT > 0 ? CT : 1)>(fr, eye_subspace, seed, pick);
        if (!done) eye_sample_serial(fr, eye_subspace, C, seed, pick);
        reinterpret_cast<uint32_t*>(a.res + pix)[3] = seed;
#pragma unroll
        for (int j = 0; j < (CT > 0 ? CT : 16); j++) {
            if (j >= C) break;
            int lv = pick[j].lv;
            const spc_vertex* L = fr.p.sampler.LVC + max(lv, 0);
            const float3 lp = f3(L->position.x, L->position.y, L->position.z);
            // endpoints facing away from each other: the contribution is exactly zero, no shadow ray and no evaluation (shade.cuh)
            if (lv >= 0 && connection_is_dead(pos, nrm, lp, f3(L->normal.x, L->normal.y, L->normal.z), L->isOrigin != 0)) lv = -1;
            a.conn_lvc[(size_t)i * C + j] = lv;
            if (lv < 0) {   // empty subspace or a connection that cannot contribute: an empty-interval shadow ray
                a.shadow[2 * ((size_t)i * C + j)] = make_float4(0.f, 0.f, 0.f, 1.0f);
                a.shadow[2 * ((size_t)i * C + j) + 1] = make_float4(0.f, 0.f, 1.f, -1.0f);
                continue;
            }
            // visibilityTest (cuProg.h:489-502 -> :463-487)
            const float3 bias_pos = lp - pos;
            const float len = length(bias_pos);
            const float3 dir = bias_pos / len;
            a.shadow[2 * ((size_t)i * C + j)] = make_float4(pos.x, pos.y, pos.z, SPC_SCENE_EPS);
            a.shadow[2 * ((size_t)i * C + j) + 1] = make_float4(dir.x, dir.y, dir.z, len - SPC_SCENE_EPS);
            a.conn_pmf[(size_t)i * C + j] = pick[j].pmf;
            n_shadow++;
        }
    }
    n_shadow = __reduce_add_sync(0xffffffffu, n_shadow);   // every lane reaches this point
    if ((threadIdx.x & 31) == 0 && n_shadow) atomicAdd(a.stat + 0, (unsigned long long)n_shadow);
}

// eye-tree label of every valid LVC vertex (the `jump_buffer` lists them), once per frame: tracing_weight_light would
// otherwise walk the eye tree for the light vertex of every single connection (shade.cuh, "cross labels")
__global__ void k_lvc_xlabel(const DevFrame fr, int n_valid, short* __restrict__ xlabel) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_valid) return;
    const int lv = fr.p.sampler.jump_buffer[j];
    const spc_vertex* L = fr.p.sampler.LVC + lv;
    xlabel[lv] = (short)eye_tree_label(fr, f3(L->position.x, L->position.y, L->position.z), f3(L->normal.x, L->normal.y, L->normal.z));
}

// connectVertex_SPCBPT (raygen.cu:253-303) for every visible connection; one lane per connection.
// Only a fraction of the slots carries a visible connection (empty slots, occluded shadow rays), so each warp first compacts
// the slots that need work into a small shared-memory queue and evaluates them 32 at a time: measured 10 of 32 lanes active per
// issued instruction without the queue (profiles/r1d_summary.md).  Every slot's term is written to its own place, so the order
// in which connections are evaluated does not affect the result (k_eye_gather sums in slot order).
__device__ __forceinline__ void eye_connect_one(const DevFrame& fr, const EyeArgs& a, int64_t k, int C) {
    const int lv = a.conn_lvc[k];
    const int pix = a.queue_cur[k / C];
    const Vtx eye = vtx_load(a.ev + pix);
    const Vtx light = vtx_load(fr.p.sampler.LVC + lv);
    const float3 c = connect_vertices(fr, eye, light, nullptr, (int)a.xlab[pix], fr.lvc_xlabel ? (int)__ldg(fr.lvc_xlabel + lv) : -1);
    const float3 res = c / a.conn_pmf[k];
    float3 term = f3(0.f);
    if (!invalid3(res)) term = res / (float)C;
    a.contrib[k] = make_float4(term.x, term.y, term.z, 0.f);
}

// Two queues per warp: connections to emitter points (light-path depth 0: one-sided emission, connection_lightSource) and to
// surface light vertices (second BSDF, general_connection) take different branches throughout connect_vertices; evaluated from one
// mixed queue a warp ran both sides for most batches (ncu: 18 of 32 lanes active per instruction, profiles/r1e_summary.md).
__global__ void __launch_bounds__(128) k_eye_connect(const DevFrame fr, const EyeArgs a) {
    __shared__ int64_t s_queue[4][2][64];
    const int C = fr.connections;
    const int64_t n = (int64_t)a.counts[a.bounce] * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int pending[2] = {0, 0};   // warp-uniform
    unsigned n_visible = 0;    // warp-uniform
    bool last = false;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + warp * 32;; base += (int64_t)gridDim.x * blockDim.x) {
        if (base >= n) last = true;   // one extra round that drains what is left
        if (!last) {
            const int64_t k = base + lane;
            int kind = -1;
            if (k < n) {
                const int lv = a.conn_lvc[k];
                if (lv >= 0 && a.visible[k]) kind = fr.p.sampler.LVC[lv].depth == 0 ? 0 : 1;
                else a.contrib[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const unsigned m = __ballot_sync(0xffffffffu, kind == t);
                if (kind == t) s_queue[warp][t][pending[t] + __popc(m & ((1u << lane) - 1u))] = k;
                pending[t] += __popc(m);
                n_visible += __popc(m);
            }
            __syncwarp();
        }
        const int threshold = last ? 1 : 32;
#pragma unroll 1
        for (;;) {
            const int t = pending[1] >= threshold ? 1 : (pending[0] >= threshold ? 0 : -1);
            if (t < 0) break;
            const int cnt = min(32, pending[t]);
            // newest `cnt` entries: the older remainder stays at the front of the queue
            if (lane < cnt) eye_connect_one(fr, a, s_queue[warp][t][pending[t] - cnt + lane], C);
            __syncwarp();
            pending[t] -= cnt;
        }
        if (last) break;
    }
    if (lane == 0 && n_visible) atomicAdd(a.stat + 1, (unsigned long long)n_visible);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Tail of the eye pass.  After a few bounces a frame has a few thousand live paths left, yet every further bounce of the wavefront
// costs six kernel launches whose duration is the latency of ONE path's dependent loads (~250 us per bounce, ~18 bounces on the
// shipped scene: half of a sequential frame).  Once the live-path count has dropped below `tail_threshold` (default 131072: house scene,
// sequential 1080p frames 10.4 ms without the tail, 9.7 / 9.5 / 9.3 ms with 32 k / 128 k / 512 k, profiles/r2e_summary.md) the remaining bounces of
// every surviving path run to completion in this one kernel, one lane per path: closest hit -> surface program -> classification ->
// C two-stage samples -> shadow rays -> connections, the loop body of __raygen__SPCBPT (raygen.cu:357-421) for that path.
// Same draws in the same order and the same fp32 accumulation order as the wavefront stages (emitter term, then the connection
// terms in connection order), so frames are bit-identical with and without the tail (tests/test_pipeline_gpu.py).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int kTailBlock = 128;
template <int CT>
__global__ void __launch_bounds__(kTailBlock) k_eye_tail(const DevFrame fr, const EyeArgs a, int first_bounce) {
    __shared__ uint2 s_stack[kSmStack * kTailBlock];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);
    const int n = a.counts[first_bounce];
    const int C = CT > 0 ? CT : fr.connections;
    unsigned n_closest = 0, n_shadow = 0, n_visible = 0, cn = 0, ct = 0;
    uint2* stack = s_stack + threadIdx.x;
    for (int i = blockIdx.x * kTailBlock + threadIdx.x; i < n; i += gridDim.x * kTailBlock) {
        const int pix = a.queue_cur[i];
        const float4 ro4 = a.rays_cur[2 * (size_t)i], rd4 = a.rays_cur[2 * (size_t)i + 1];
        float3 ray_origin = f3(ro4.x, ro4.y, ro4.z), ray_direction = f3(rd4.x, rd4.y, rd4.z);
        Vtx last = vtx_load(a.ev + pix);
        float4 pre = a.pre[pix];
        float4 res = a.res[pix];
        uint32_t seed = __float_as_uint(res.w);
        int last_x = (int)a.xlab[pix];
        for (int bounce = first_bounce; bounce <= fr.max_depth; bounce++) {
            const TravRay r{ray_origin.x, ray_origin.y, ray_origin.z, ray_direction.x, ray_direction.y, ray_direction.z, SPC_SCENE_EPS, 1e16f};
            TravHit h;
            n_closest++;
            if (!traverse_bvh8<false, false>(fr.sc.nodes, fr.sc.tris, r, true, stack, kTailBlock, h, cn, ct, s_lut)) break;   // miss
            const LocalGeom g = hit_geometry(fr.sc, h.prim, h.u, h.v);
            Vtx mid;
            if (g.light >= 0) {
                if (eye_hits_light(fr, last, f3(pre.x, pre.y, pre.z), pre.w, g, h.t, ray_direction, mid, last_x)) {
                    const float3 ans = mid.flux / mid.pdf / mid.RMIS_pointer;
                    if (!invalid3(ans)) { res.x += ans.x; res.y += ans.y; res.z += ans.z; }
                }
                break;
            }
            SurfaceOut so;
            surface_hit(fr, last, f3(pre.x, pre.y, pre.z), pre.w, g, h.t, ray_direction, false, seed, mid, so, last_x, true);
            int eye_subspace, cross;
            if (fr.eye_ctree && fr.light_ctree) ctree_label2(fr.eye_ctree, fr.light_ctree, mid.position, mid.normal, eye_subspace, cross);
            else tree_label2(fr.p.subspace_info.eye_tree, fr.p.subspace_info.light_tree, mid.position, mid.normal, eye_subspace, cross);
            mid.subspaceId = (short)eye_subspace;
            ConnPick pick[CT > 0 ? CT : 16];
            bool sampled = false;
            if (CT > 0) sampled = eye_sample_lockstep<(CT > 0 ? CT : 1)>(fr, eye_subspace, seed, pick);
            if (!sampled) eye_sample_serial(fr, eye_subspace, C, seed, pick);
#pragma unroll 1
            for (int j = 0; j < C; j++) {
                const int lv = pick[j].lv;
                if (lv < 0) continue;
                const Vtx light = vtx_load(fr.p.sampler.LVC + lv);
                if (connection_is_dead(mid.position, mid.normal, light.position, light.normal, light.isOrigin != 0)) continue;   // contributes +0
                // visibilityTest (cuProg.h:489-502 -> :463-487)
                const float3 bias_pos = light.position - mid.position;
                const float len = length(bias_pos);
                const float3 dir = bias_pos / len;
                n_shadow++;
                const float tmax = len - SPC_SCENE_EPS;
                bool visible = true;
                if (tmax > SPC_SCENE_EPS) {   // an empty interval hits nothing (k_trace_persist skips it the same way)
                    const TravRay sr{mid.position.x, mid.position.y, mid.position.z, dir.x, dir.y, dir.z, SPC_SCENE_EPS, tmax};
                    TravHit sh;
                    visible = !traverse_bvh8<true, false>(fr.sc.nodes, fr.sc.tris, sr, false, stack, kTailBlock, sh, cn, ct, s_lut);
                }
                if (!visible) continue;
                n_visible++;
                const float3 c = connect_vertices(fr, mid, light, nullptr, cross, fr.lvc_xlabel ? (int)__ldg(fr.lvc_xlabel + lv) : -1);
                const float3 rr = c / pick[j].pmf;
                if (!invalid3(rr)) {
                    const float3 term = rr / (float)C;
                    res.x += term.x; res.y += term.y; res.z += term.z;   // connection order, as k_eye_gather
                }
            }
            if (so.done) break;
            last = mid;
            last_x = cross;
            pre = make_float4(so.next_flux.x, so.next_flux.y, so.next_flux.z, so.next_singlePdf);
            ray_origin = g.P;
            ray_direction = so.dir;
        }
        res.w = __uint_as_float(seed);
        a.res[pix] = res;
    }
    // work counters (spc_eye_stats_get): [2] closest-hit rays of the tail
    n_closest = __reduce_add_sync(0xffffffffu, n_closest);
    n_shadow = __reduce_add_sync(0xffffffffu, n_shadow);
    n_visible = __reduce_add_sync(0xffffffffu, n_visible);
    if ((threadIdx.x & 31) == 0) {
        if (n_closest) atomicAdd(a.stat + 2, (unsigned long long)n_closest);
        if (n_shadow) atomicAdd(a.stat + 0, (unsigned long long)n_shadow);
        if (n_visible) atomicAdd(a.stat + 1, (unsigned long long)n_visible);
    }
}

// result += res / CONNECTION_N, in connection order (raygen.cu:415)
__global__ void k_eye_gather(const DevFrame fr, const EyeArgs a) {
    const int C = fr.connections;
    const int n = a.counts[a.bounce];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int pix = a.queue_cur[i];
        float4 r = a.res[pix];
        bool any = false;
        for (int j = 0; j < C; j++) {
            if (a.conn_lvc[(size_t)i * C + j] < 0) continue;
            const float4 t = a.contrib[(size_t)i * C + j];
            r.x += t.x; r.y += t.y; r.z += t.z;
            any = true;
        }
        if (any) a.res[pix] = r;
    }
}

// running mean + ToneMap (raygen.cu:50-58, :430-442) + make_color (src/cuda/helpers.h:35-67)
__device__ __forceinline__ unsigned quantize8(float x) {
    x = clampf(x, 0.0f, 1.0f);
    return min((unsigned)(x * 256.0f), 255u);
}
__device__ __forceinline__ float to_srgb(float c) {
    const float invGamma = 1.0f / 2.4f;
    const float powed = cm_powf(c, invGamma);
    return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f;
}
__global__ void k_accumulate(const DevFrame fr, const float4* __restrict__ res, int n_pix, const int* __restrict__ pixels /* null: all pixels */) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pix) return;
    const int i = pixels ? pixels[k] : k;
    const float4 r = res[i];
    float3 c = f3(r.x, r.y, r.z);
    if (fr.p.subframe_index > 0) {
        const float t = 1.0f / (float)(int)(fr.p.subframe_index + 1);
        const spc_float4 prev = fr.p.accum_buffer[i];
        c = lerp3(f3(prev.x, prev.y, prev.z), c, t);
    }
    fr.p.accum_buffer[i] = spc_float4{c.x, c.y, c.z, 1.0f};
    if (fr.p.frame_buffer) {
        const float lum = 0.3f * c.x + 0.6f * c.y + 0.1f * c.z;
        const float s = 1.0f + 1 * lum / 1.5f;
        const float inv = 1.0f / s;   // `c * 1.0f / s` on a float4 multiplies by the reciprocal (sutil/vec_math.h:720-724)
        const float3 v = f3(clampf(c.x * 1.0f * inv, 0.f, 1.f), clampf(c.y * 1.0f * inv, 0.f, 1.f), clampf(c.z * 1.0f * inv, 0.f, 1.f));
        fr.p.frame_buffer[i] = quantize8(to_srgb(v.x)) | (quantize8(to_srgb(v.y)) << 8) | (quantize8(to_srgb(v.z)) << 16) | (255u << 24);
    }
}

// Read-out of a sample-partitioned render: out = sum_k w_k * accum_k in the order given (fp32, one multiply-add chain per
// channel, no contraction), then the display transform of k_accumulate.  Used for frame lanes on one GPU (host/spcbpt_main.cpp
// --lanes) and after the NCCL gather of per-rank buffers.
__global__ void k_merge_accum(const spc_float4* const* __restrict__ bufs, const float* __restrict__ w, int n, int n_pix,
                              spc_float4* __restrict__ out, uint32_t* __restrict__ frame) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    float3 c = f3(0.f, 0.f, 0.f);
    for (int k = 0; k < n; k++) {
        const spc_float4 a = bufs[k][i];
        const float wk = w[k];
        c = f3(c.x + wk * a.x, c.y + wk * a.y, c.z + wk * a.z);
    }
    if (out) out[i] = spc_float4{c.x, c.y, c.z, 1.0f};
    if (frame) {
        const float lum = 0.3f * c.x + 0.6f * c.y + 0.1f * c.z;
        const float inv = 1.0f / (1.0f + 1 * lum / 1.5f);
        const float3 v = f3(clampf(c.x * 1.0f * inv, 0.f, 1.f), clampf(c.y * 1.0f * inv, 0.f, 1.f), clampf(c.z * 1.0f * inv, 0.f, 1.f));
        frame[i] = quantize8(to_srgb(v.x)) | (quantize8(to_srgb(v.y)) << 8) | (quantize8(to_srgb(v.z)) << 16) | (255u << 24);
    }
}

void merge_accum(Context& c, const spc_float4* const* bufs_host, const float* weights_host, int n, int n_pix, spc_float4* out, uint32_t* frame) {
    SPC_REQUIRE(bufs_host && weights_host && n > 0 && n <= 64 && n_pix > 0 && (out || frame), SPC_ERR_INVALID, "spc_merge_accum: bad arguments");
    c.merge_ptrs.alloc(64);
    c.merge_w.alloc(64);
    SPC_CUDA(cudaMemcpyAsync(c.merge_ptrs.p, bufs_host, n * sizeof(void*), cudaMemcpyHostToDevice, c.stream));
    SPC_CUDA(cudaMemcpyAsync(c.merge_w.p, weights_host, n * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    k_merge_accum<<<(n_pix + 255) / 256, 256, 0, c.stream>>>((const spc_float4* const*)c.merge_ptrs.p, c.merge_w.p, n, n_pix, out, frame);
    SPC_CUDA(cudaGetLastError());
    SPC_CUDA(cudaStreamSynchronize(c.stream));   // the host arrays may go away after the call
    c.launches++;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Sorting between bounces (spc_set_option "sort_hits").  From the second bounce on the hit points of a wavefront are scattered over
// the scene although the queue is still in (compacted) pixel order.  After the closest-hit pass of bounce b >= 1 the queue entries
// {pixel, ray, hit} are re-ordered by the Morton code of their hit point (5 bits per axis of the scene box, misses last): every
// later stage of the bounce then works on spatially coherent entries -- material / texture fetches in k_eye_shade, the eye-subspace
// row of the Gamma CDF in k_eye_sample, shadow-ray origins, and the next bounce's rays (their origins ARE these hit points).
// A counting sort over 2^15 + 1 bins: keys + histogram, one-block scan, scatter.  The order inside a bin is whatever the atomics
// give: nothing depends on queue order (every result is written per pixel or per queue slot and summed in connection order), so
// frames stay bit-identical with and without the sort (tests/test_pipeline_gpu.py).
// ---------------------------------------------------------------------------------------------------------------------------
constexpr int kSortBits = 15, kSortBins = (1 << kSortBits) + 1;
__device__ __forceinline__ uint32_t morton_spread5(uint32_t v) {   // 5 bits -> every third bit
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6) | ((v & 16u) << 8);
}
__global__ void k_sort_keys(const float4* __restrict__ rays, const float4* __restrict__ hits, const int* __restrict__ n_dev, float3 lo, float3 inv_extent,
                            uint32_t* __restrict__ keys, int* __restrict__ hist) {
    const int n = *n_dev;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 h = hits[i];
        uint32_t key = kSortBins - 1;   // miss
        if (__float_as_int(h.w) >= 0) {
            const float4 ro = rays[2 * (size_t)i], rd = rays[2 * (size_t)i + 1];
            const float px = (ro.x + h.x * rd.x - lo.x) * inv_extent.x, py = (ro.y + h.x * rd.y - lo.y) * inv_extent.y, pz = (ro.z + h.x * rd.z - lo.z) * inv_extent.z;
            const uint32_t qx = (uint32_t)min(31, max(0, (int)(px * 32.f))), qy = (uint32_t)min(31, max(0, (int)(py * 32.f))), qz = (uint32_t)min(31, max(0, (int)(pz * 32.f)));
            key = morton_spread5(qx) | (morton_spread5(qy) << 1) | (morton_spread5(qz) << 2);
        }
        keys[i] = key;
        atomicAdd(hist + key, 1);
    }
}
__global__ void k_sort_scan(int* __restrict__ hist) {   // exclusive scan of kSortBins counters, one block of 1024
    __shared__ int s_part[1024];
    constexpr int per = (kSortBins + 1023) / 1024;
    const int lo = min(kSortBins, (int)threadIdx.x * per), hi = min(kSortBins, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += hist[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {   // Hillis-Steele inclusive scan
        const int v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = s_part[threadIdx.x] - sum;
    for (int i = lo; i < hi; i++) {
        const int v = hist[i];
        hist[i] = run;
        run += v;
    }
}
__global__ void k_sort_scatter(const uint32_t* __restrict__ keys, int* __restrict__ offsets, const int* __restrict__ n_dev, const float4* __restrict__ rays,
                               const float4* __restrict__ hits, const int* __restrict__ queue, float4* __restrict__ rays_out, float4* __restrict__ hits_out,
                               int* __restrict__ queue_out) {
    const int n = *n_dev;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int dst = atomicAdd(offsets + keys[i], 1);
        rays_out[2 * (size_t)dst] = rays[2 * (size_t)i];
        rays_out[2 * (size_t)dst + 1] = rays[2 * (size_t)i + 1];
        hits_out[dst] = hits[i];
        queue_out[dst] = queue[i];
    }
}

void launch_eye_pass(Context& c, int width, int height) {
    NvtxRange range("spc: SPCBPT_eye");
    SPC_REQUIRE(c.has_params, SPC_ERR_INVALID, "spc_launch: spc_set_params has not been called");
    SPC_REQUIRE(width > 0 && height > 0 && (unsigned)width == c.params.width && (unsigned)height == c.params.height, SPC_ERR_INVALID,
                "spc_launch(SPCBPT_eye): launch size %dx%d differs from MyParams %ux%u", width, height, c.params.width, c.params.height);
    SPC_REQUIRE(c.params.accum_buffer, SPC_ERR_INVALID, "spc_launch(SPCBPT_eye): MyParams::accum_buffer is null");
    const spc_subspace_sampler& S = c.params.sampler;
    SPC_REQUIRE(S.LVC && S.subspace && S.cmfs && S.jump_buffer, SPC_ERR_INVALID, "spc_launch(SPCBPT_eye): MyParams::sampler is not set (run LVC_Process first)");
    SPC_REQUIRE(!c.params.subspace_info.light_tree || c.params.subspace_info.CMFGamma, SPC_ERR_INVALID,
                "spc_launch(SPCBPT_eye): light_tree without CMFGamma");
    const size_t P = (size_t)width * height;
    SPC_REQUIRE(P < 0x7fffffffull / 16, SPC_ERR_INVALID, "spc_launch: image too large");
    DevFrame fr = make_dev_frame(c);
    const int C = c.connections;
    EyeBuffers& e = c.eye;
    if (e.pixels < P || e.conns != C) {
        e.ev.alloc(P); e.pre.alloc(P); e.res.alloc(P); e.xlab.alloc(P);
        for (int k = 0; k < 2; k++) { e.rays[k].alloc(P); e.queue[k].alloc(P); }
        e.hits.alloc(P);
        e.rays_sorted.alloc(P); e.queue_sorted.alloc(P); e.hits_sorted.alloc(P); e.sort_keys.alloc(P); e.sort_hist.alloc(kSortBins);
        e.shadow.alloc(P * C); e.visible.alloc(P * C); e.conn_lvc.alloc(P * C); e.conn_pmf.alloc(P * C); e.contrib.alloc(P * C);
        e.pixels = P; e.conns = C;
    }
    const int n_counts = fr.max_depth + 4;
    e.counts.alloc(n_counts);
    e.stat.alloc(8);
    if (!c.h_pinned) SPC_CUDA(cudaMallocHost((void**)&c.h_pinned, 64 * sizeof(int)));
    cudaStream_t st = c.stream;
    SPC_CUDA(cudaMemsetAsync(e.counts.p, 0, n_counts * sizeof(int), st));
    SPC_CUDA(cudaMemsetAsync(e.stat.p, 0, 8 * sizeof(unsigned long long), st));
    // option "stage_timing": events around every stage of every bounce (spc_eye_stats_get sums them per stage)
    const bool timed = c.opt[OPT_STAGE_TIMING] != 0;
    constexpr int kEvPerBounce = 7;
    if (timed) {
        const size_t need = (size_t)(fr.max_depth + 1) * kEvPerBounce + 2;
        while (e.stage_ev.size() < need) {
            cudaEvent_t ev;
            SPC_CUDA(cudaEventCreate(&ev));
            e.stage_ev.push_back(ev);
        }
        SPC_CUDA(cudaEventRecord(e.stage_ev[0], st));
    }
    auto mark = [&](int b, int k) {
        if (timed) SPC_CUDA(cudaEventRecord(e.stage_ev[2 + (size_t)b * kEvPerBounce + k], st));
    };
    e.last_timed = timed;
    e.last_bounces = 0;

    EyeArgs a;
    a.ev = e.ev.p; a.pre = e.pre.p; a.res = e.res.p;
    a.hits = (const float4*)e.hits.p; a.shadow = (float4*)e.shadow.p; a.visible = e.visible.p;
    a.conn_lvc = e.conn_lvc.p; a.conn_pmf = e.conn_pmf.p; a.contrib = e.contrib.p; a.counts = e.counts.p;
    a.first_prim = c.dbg_first_prim; a.first_label = c.dbg_first_label;
    a.xlab = e.xlab.p;
    a.stat = e.stat.p;
    // cross labels of the light vertices: only when the sampler is the one spc_lvc_process built (then jump_buffer indexes
    // c.lvc.n slots); a caller-made sampler keeps the in-kernel tree walk
    fr.lvc_xlabel = nullptr;
    fr.gamma_guide = c.params.subspace_info.light_tree ? gamma_guide_lookup(c.params.subspace_info.CMFGamma, c.K) : nullptr;
    fr.lvc_guide = (S.cmfs == c.lvc.cmfs.p && S.subspace == c.lvc.subspace.p && c.lvc.guide_valid) ? c.lvc.guide.p : nullptr;
    if (S.jump_buffer == c.lvc.jump.p && c.lvc.n > 0 && S.vertex_count > 0 && S.vertex_count <= c.lvc.n) {
        e.lvc_xlabel.alloc((size_t)c.lvc.n);
        k_lvc_xlabel<<<(S.vertex_count + 255) / 256, 256, 0, st>>>(fr, S.vertex_count, e.lvc_xlabel.p);
        c.launches++;
        fr.lvc_xlabel = e.lvc_xlabel.p;
    }
    a.bounce = 0;
    a.rays_cur = (float4*)e.rays[0].p; a.rays_next = (float4*)e.rays[1].p;
    a.queue_cur = e.queue[0].p; a.queue_next = e.queue[1].p;

    if (c.opt[OPT_REFERENCE_SEARCH]) {
        // test switch (spc_set_option "reference_search", tests/test_pipeline_gpu.py): the reference's bisect and the reference-layout tree walks instead of guide tables,
        // compact trees and the cached light-vertex labels -- frames must come out bit-identical either way
        fr.gamma_guide = nullptr;
        fr.lvc_guide = nullptr;
        fr.eye_ctree = nullptr;
        fr.light_ctree = nullptr;
        fr.lvc_xlabel = nullptr;
    }
    const int nP = (int)P;
    // queue of the first bounce = pixel order: an identity array filled once per allocation; per frame only counts[0] (and the
    // optional parity dumps) are initialised.  Tile partition (spc_set_tile_partition): the queue is this rank's pixel list instead.
    const bool tiled = c.tile_world > 1;
    int n_first = nP;
    if (tiled) {
        if (c.tile_key[0] != width || c.tile_key[1] != height || c.tile_key[2] != c.tile_rank || c.tile_key[3] != c.tile_world) {
            // StaticWorkDistribution::getSamplePixel (sutil/WorkDistribution.h:59-82): 8 x 4 tiles, strips of `world` tiles, the
            // rank's tile shifted by one per strip row; pixels beyond the image are dropped
            constexpr int TW = 8, TH = 4;
            const int G = c.tile_world, strip_w = TW * G;
            const int cols = (width + strip_w - 1) / strip_w, rows = (height + TH - 1) / TH;
            std::vector<int> px;
            px.reserve((size_t)P / G + 64);
            for (int sidx = 0; sidx < rows * cols * TW * TH; sidx++) {
                const int strip = sidx / (TW * TH), sy = strip / cols, sx = strip - sy * cols;
                const int tp = sidx - strip * TW * TH, ty = tp / TW, tx = tp - ty * TW;
                const int x = sx * strip_w + tx + (c.tile_rank + sy % G) % G * TW, y = sy * TH + ty;
                if (x < width && y < height) px.push_back(y * width + x);
            }
            c.tile_pixels.alloc(px.size());
            SPC_CUDA(cudaMemcpyAsync(c.tile_pixels.p, px.data(), px.size() * sizeof(int), cudaMemcpyHostToDevice, st));
            SPC_CUDA(cudaStreamSynchronize(st));
            c.n_tile_pixels = (int)px.size();
            c.tile_key[0] = width; c.tile_key[1] = height; c.tile_key[2] = c.tile_rank; c.tile_key[3] = c.tile_world;
        }
        n_first = c.n_tile_pixels;
    }
    const bool fill_ident = !tiled && e.ident_pixels != P;
    if (fill_ident) e.queue_ident.alloc(P);
    const int n_work = (fill_ident || a.first_prim || a.first_label) ? nP : 1;
    k_eye_begin<<<(n_work + 255) / 256, 256, 0, st>>>(a, nP, n_first, n_work, e.queue_ident.p, fill_ident ? 1 : 0);
    if (fill_ident) e.ident_pixels = P;
    const int* first_queue = tiled ? c.tile_pixels.p : e.queue_ident.p;
    c.launches++;
    CamGen cam;
    cam.eye = make_float3(c.params.eye.x, c.params.eye.y, c.params.eye.z);
    cam.U = make_float3(c.params.U.x, c.params.U.y, c.params.U.z);
    cam.V = make_float3(c.params.V.x, c.params.V.y, c.params.V.z);
    cam.W = make_float3(c.params.W.x, c.params.W.y, c.params.W.z);
    cam.width = c.params.width;
    cam.height = c.params.height;
    cam.sample_index = c.params.subframe_index * c.seed_stride + c.seed_offset;
    cam.pixels = tiled ? c.tile_pixels.p : nullptr;
    const int grid_cap = c.sm_count * 16;
    int64_t n_max = n_first;   // host-side upper bound on the live paths (refreshed by the lagged read-backs)
    // The size of every bounce's queue is copied to pinned memory behind the bounce, and the host looks at it kLag bounces later:
    // it stops when a queue was empty and shrinks the grids, but it never drains the stream (a full synchronisation every 4th
    // bounce left the GPU idle for a host round trip each time and made the frame time follow the host's scheduling noise).
    constexpr int kLag = 3, kRing = 8;
    const bool sort_hits = c.opt[OPT_SORT_HITS] != 0;
    const int64_t sort_min = 4096;   // below this a bounce is latency-bound anyway (and the tail kernel takes over soon)
    const float3 sort_lo = make_float3(c.geom.scene_lo[0], c.geom.scene_lo[1], c.geom.scene_lo[2]);
    const float3 sort_inv = make_float3(1.0f / std::max(1e-30f, c.geom.scene_hi[0] - c.geom.scene_lo[0]), 1.0f / std::max(1e-30f, c.geom.scene_hi[1] - c.geom.scene_lo[1]),
                               1.0f / std::max(1e-30f, c.geom.scene_hi[2] - c.geom.scene_lo[2]));
    const int64_t tail_threshold = c.opt[OPT_TAIL_THRESHOLD] < 0 ? 0 : (c.opt[OPT_TAIL_THRESHOLD] == 0 ? 131072 : c.opt[OPT_TAIL_THRESHOLD]);
    if (!c.eye_events[0])
        for (int k = 0; k < kRing; k++)
            SPC_CUDA(cudaEventCreateWithFlags(&c.eye_events[k], cudaEventDisableTiming | (c.opt[OPT_BLOCKING_SYNC] ? cudaEventBlockingSync : 0)));
    int* h_ring = c.h_pinned + 16;
    // loop of raygen.cu:357-421: a path is traced while !done && depth <= max_depth, i.e. bounces 0..max_depth
    for (int b = 0; b <= fr.max_depth; b++) {
        a.bounce = b;
        a.rays_cur = (float4*)e.rays[b & 1].p; a.rays_next = (float4*)e.rays[(b + 1) & 1].p;
        a.queue_cur = b == 0 ? const_cast<int*>(first_queue) : e.queue[b & 1].p; a.queue_next = e.queue[(b + 1) & 1].p;
        mark(b, 0);
        nvtxRangePushA("eye: closest hits");
        if (b == 0) launch_trace_closest_camera(c, cam, n_first, SPC_RAYFLAG_CULL_BACK_FACING, e.hits.p);   // generates the camera rays itself
        else launch_trace_closest_q(c, (const spc_ray*)a.rays_cur, e.counts.p + b, 1, n_max, SPC_RAYFLAG_CULL_BACK_FACING, e.hits.p);
        if (sort_hits && b >= 1 && n_max >= sort_min) {
            // re-order the queue by hit-point Morton code (see k_sort_keys): the rest of the bounce reads the sorted copies
            const int gs = (int)std::min<int64_t>((n_max + 255) / 256, grid_cap);
            SPC_CUDA(cudaMemsetAsync(e.sort_hist.p, 0, kSortBins * sizeof(int), st));
            k_sort_keys<<<gs, 256, 0, st>>>(a.rays_cur, (const float4*)e.hits.p, e.counts.p + b, sort_lo, sort_inv, e.sort_keys.p, e.sort_hist.p);
            k_sort_scan<<<1, 1024, 0, st>>>(e.sort_hist.p);
            k_sort_scatter<<<gs, 256, 0, st>>>(e.sort_keys.p, e.sort_hist.p, e.counts.p + b, a.rays_cur, (const float4*)e.hits.p, a.queue_cur, (float4*)e.rays_sorted.p,
                                               (float4*)e.hits_sorted.p, e.queue_sorted.p);
            c.launches += 3;
            a.rays_cur = (float4*)e.rays_sorted.p;
            a.queue_cur = e.queue_sorted.p;
            a.hits = (const float4*)e.hits_sorted.p;
        } else {
            a.hits = (const float4*)e.hits.p;
        }
        nvtxRangePop();
        mark(b, 1);
        nvtxRangePushA("eye: shade + sample");
        const int g1 = (int)std::min<int64_t>((n_max + 127) / 128, grid_cap);
        k_eye_shade<<<g1, 128, 0, st>>>(fr, a);
        mark(b, 2);
        switch (C) {
            case 1: k_eye_sample<1><<<g1, 128, 0, st>>>(fr, a); break;
            case 2: k_eye_sample<2><<<g1, 128, 0, st>>>(fr, a); break;
            case 3: k_eye_sample<3><<<g1, 128, 0, st>>>(fr, a); break;
            case 4: k_eye_sample<4><<<g1, 128, 0, st>>>(fr, a); break;
            default: k_eye_sample<0><<<g1, 128, 0, st>>>(fr, a); break;
        }
        nvtxRangePop();
        mark(b, 3);
        nvtxRangePushA("eye: shadow rays");
        launch_trace_occlusion_q(c, (const spc_ray*)a.shadow, e.counts.p + b, C, n_max * C, e.visible.p);
        nvtxRangePop();
        mark(b, 4);
        nvtxRangePushA("eye: connect + gather");
        const int g2 = (int)std::min<int64_t>((n_max * C + 127) / 128, grid_cap);
        k_eye_connect<<<g2, 128, 0, st>>>(fr, a);
        mark(b, 5);
        k_eye_gather<<<g1, 128, 0, st>>>(fr, a);
        nvtxRangePop();
        mark(b, 6);
        e.last_bounces = b + 1;
        c.launches += 4;
        SPC_CUDA(cudaGetLastError());
        if (b < fr.max_depth) {
            SPC_CUDA(cudaMemcpyAsync(h_ring + (b % kRing), e.counts.p + b + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            SPC_CUDA(cudaEventRecord(c.eye_events[b % kRing], st));
            if (b >= kLag) {
                const int pb = b - kLag;   // live paths entering bounce pb + 1: an upper bound for every later bounce
                SPC_CUDA(cudaEventSynchronize(c.eye_events[pb % kRing]));
                n_max = h_ring[pb % kRing];
                if (n_max == 0) break;
                if (tail_threshold > 0 && n_max <= tail_threshold && b + 1 <= fr.max_depth) {
                    NvtxRange tail_range("eye: tail kernel");
                    // few paths left: the remaining bounces of every survivor in one kernel (queue of bounce b + 1 = the `next` buffers)
                    a.bounce = b + 1;
                    a.rays_cur = (float4*)e.rays[(b + 1) & 1].p;
                    a.queue_cur = e.queue[(b + 1) & 1].p;
                    const int gt = (int)std::min<int64_t>((n_max + kTailBlock - 1) / kTailBlock, grid_cap);
                    switch (C) {
                        case 1: k_eye_tail<1><<<gt, kTailBlock, 0, st>>>(fr, a, b + 1); break;
                        case 2: k_eye_tail<2><<<gt, kTailBlock, 0, st>>>(fr, a, b + 1); break;
                        case 3: k_eye_tail<3><<<gt, kTailBlock, 0, st>>>(fr, a, b + 1); break;
                        case 4: k_eye_tail<4><<<gt, kTailBlock, 0, st>>>(fr, a, b + 1); break;
                        default: k_eye_tail<0><<<gt, kTailBlock, 0, st>>>(fr, a, b + 1); break;
                    }
                    c.launches++;
                    SPC_CUDA(cudaGetLastError());
                    break;
                }
            }
        }
    }
    k_accumulate<<<(n_first + 255) / 256, 256, 0, st>>>(fr, e.res.p, n_first, tiled ? c.tile_pixels.p : nullptr);
    c.launches++;
    if (timed) SPC_CUDA(cudaEventRecord(e.stage_ev[1], st));
    SPC_CUDA(cudaGetLastError());
}

// spc_eye_stats_get: counters (and, under option "stage_timing", per-stage device times) of the last eye pass
void eye_stats(Context& c, spc_eye_stats* out) {
    EyeBuffers& e = c.eye;
    *out = spc_eye_stats{};
    SPC_REQUIRE(e.last_bounces > 0 && e.stat.p && e.counts.p, SPC_ERR_INVALID, "spc_eye_stats_get: no eye pass has run on this context");
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    std::vector<int> counts((size_t)e.last_bounces);
    unsigned long long stat[8];
    SPC_CUDA(cudaMemcpy(counts.data(), e.counts.p, counts.size() * sizeof(int), cudaMemcpyDeviceToHost));
    SPC_CUDA(cudaMemcpy(stat, e.stat.p, sizeof(stat), cudaMemcpyDeviceToHost));
    out->bounces = e.last_bounces;
    for (int v : counts) out->closest_rays += (uint64_t)v;
    out->shadow_slots = out->closest_rays * (uint64_t)c.connections;   // slots the occlusion KERNEL scanned (the tail kernel traces its own)
    out->closest_rays += stat[2];                                       // rays of the tail kernel
    out->shadow_rays = stat[0];
    out->visible_connections = stat[1];
    if (e.last_timed) {
        constexpr int kEvPerBounce = 7;
        static const int stage_of[6] = {SPC_STAGE_TRACE, SPC_STAGE_SHADE, SPC_STAGE_SAMPLE, SPC_STAGE_SHADOW, SPC_STAGE_CONNECT, SPC_STAGE_GATHER};
        float sum = 0.f;
        for (int b = 0; b < e.last_bounces; b++)
            for (int k = 0; k < 6; k++) {
                float ms = 0.f;
                SPC_CUDA(cudaEventElapsedTime(&ms, e.stage_ev[2 + (size_t)b * kEvPerBounce + k], e.stage_ev[2 + (size_t)b * kEvPerBounce + k + 1]));
                out->stage_ms[stage_of[k]] += ms;
                sum += ms;
            }
        float total = 0.f;
        SPC_CUDA(cudaEventElapsedTime(&total, e.stage_ev[0], e.stage_ev[1]));
        out->stage_ms[SPC_STAGE_TOTAL] = total;
        out->stage_ms[SPC_STAGE_OTHER] = total - sum;
        out->timed = 1;
    }
}

}  // namespace spc
