// api.cu -- the extern "C" surface of libspcbpt_b200.so (declared in include/spcbpt_b200.h).
#include <cstdarg>
#include <cstring>
#include "common.cuh"

namespace spc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace spc

using spc::Context;

#define SPC_API_BEGIN                                                                            \
    if (!ctx) {                                                                                  \
        spc::set_error("null context");                                                          \
        return SPC_ERR_INVALID;                                                                  \
    }                                                                                            \
    Context& c = ctx->c;                                                                         \
    (void)c;                                                                                     \
    try {                                                                                        \
        cudaSetDevice(c.device);

#define SPC_API_END                                                                              \
    }                                                                                            \
    catch (const spc::CudaFailure& f) { return f.code; }                                         \
    catch (const std::exception& e) {                                                            \
        spc::set_error("exception: %s", e.what());                                               \
        return SPC_ERR_INVALID;                                                                  \
    }                                                                                            \
    return SPC_OK;

// Host-buffer batches are PCIe-bound (32 B per ray in, 16 B / 1 B out against ~0.4 ns of traversal per ray), so they run as a
// three-stream pipeline over chunks of 2^21 rays: the H2D copy of chunk i+1, the traversal of chunk i and the D2H copy of chunk
// i-1 overlap (PCIe is full duplex).  Same kernels and the same per-ray results as the device-buffer calls.
namespace {
constexpr int64_t kPipeChunk = 1 << 21;
constexpr int kPipeStreams = 3;

template <class Launch, class CopyBack>
void pipelined_host_batch(Context& c, const spc_ray* rays_host, int64_t n, Launch launch, CopyBack copy_back) {
    if (!c.pipe_streams[0]) {
        for (int k = 0; k < kPipeStreams; k++) SPC_CUDA(cudaStreamCreateWithFlags(&c.pipe_streams[k], cudaStreamNonBlocking));
        SPC_CUDA(cudaEventCreateWithFlags(&c.pipe_event, cudaEventDisableTiming));
    }
    cudaStream_t user = c.stream;
    SPC_CUDA(cudaEventRecord(c.pipe_event, user));   // everything enqueued so far (scene upload, ...) comes first
    for (int k = 0; k < kPipeStreams; k++) SPC_CUDA(cudaStreamWaitEvent(c.pipe_streams[k], c.pipe_event, 0));
    try {
        int k = 0;
        for (int64_t off = 0; off < n; off += kPipeChunk, k = (k + 1) % kPipeStreams) {
            const int64_t m = n - off < kPipeChunk ? n - off : kPipeChunk;
            c.stream = c.pipe_streams[k];
            SPC_CUDA(cudaMemcpyAsync(c.scratch_rays.p + off, rays_host + off, m * sizeof(spc_ray), cudaMemcpyHostToDevice, c.stream));
            launch(off, m);
            copy_back(off, m, c.stream);
        }
    } catch (...) {
        // copies to/from the caller's host buffers may still be in flight: drain them before the error reaches the caller
        c.stream = user;
        for (int k = 0; k < kPipeStreams; k++) cudaStreamSynchronize(c.pipe_streams[k]);
        throw;
    }
    c.stream = user;
    for (int k = 0; k < kPipeStreams; k++) SPC_CUDA(cudaStreamSynchronize(c.pipe_streams[k]));
}
}  // namespace

extern "C" {

const char* spc_last_error(void) { return spc::g_err; }
const char* spc_version(void) { return "spcbpt_b200 0.1 (sm_100a)"; }

int spc_create(int device, int K, int K_light, int connections, spc_context** out) {
    if (!out) {
        spc::set_error("spc_create: out is null");
        return SPC_ERR_INVALID;
    }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        spc::set_error("spc_create: no CUDA device (%s); this library has no CPU fallback",
                       e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return SPC_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) {
        spc::set_error("spc_create: device %d out of range [0,%d)", device, ndev);
        return SPC_ERR_INVALID;
    }
    if (K == 0) K = 1000;
    if (K_light == 0) K_light = int(0.2 * K);
    if (connections == 0) connections = 3;
    if (K < 2 || K > 32767 || K_light < 1 || K_light >= K || connections < 1 || connections > 16) {
        spc::set_error("spc_create: bad K=%d K_light=%d connections=%d (subspace ids are int16)", K, K_light, connections);
        return SPC_ERR_INVALID;
    }
    spc_context* ctx = new spc_context();
    ctx->c.device = device;
    ctx->c.K = K;
    ctx->c.K_light = K_light;
    ctx->c.connections = connections;
    try {
        SPC_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        SPC_CUDA(cudaGetDeviceProperties(&prop, device));
        ctx->c.sm_count = prop.multiProcessorCount;
        ctx->c.counters.alloc(8);
    } catch (const spc::CudaFailure& f) {
        delete ctx;
        return f.code;
    }
    *out = ctx;
    return SPC_OK;
}

void spc_destroy(spc_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->c.device);
    if (ctx->c.h_pinned) cudaFreeHost(ctx->c.h_pinned);
    for (cudaEvent_t ev : ctx->c.eye_events)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->c.eye.stage_ev) cudaEventDestroy(ev);
    for (cudaStream_t st : ctx->c.pipe_streams)
        if (st) cudaStreamDestroy(st);
    if (ctx->c.pipe_event) cudaEventDestroy(ctx->c.pipe_event);
    spc::gamma_guide_forget(&ctx->c, true);
    spc_comm_destroy(ctx);
    delete ctx;
}

int spc_set_stream(spc_context* ctx, void* cuda_stream) {
    SPC_API_BEGIN
    c.stream = (cudaStream_t)cuda_stream;
    SPC_API_END
}

int spc_synchronize(spc_context* ctx) {
    SPC_API_BEGIN
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    SPC_API_END
}

int64_t spc_launch_count(spc_context* ctx) { return ctx ? ctx->c.launches : 0; }

int spc_scene_upload(spc_context* ctx, const spc_mesh* meshes, int n_meshes, const spc_pbr* materials,
                     int n_materials, const spc_light* lights, int n_lights, const spc_texture* textures,
                     int n_textures) {
    SPC_API_BEGIN
    SPC_REQUIRE(meshes && n_meshes > 0, SPC_ERR_INVALID, "spc_scene_upload: no meshes");
    SPC_REQUIRE(n_materials >= 0 && n_lights >= 0 && n_textures >= 0, SPC_ERR_INVALID, "spc_scene_upload: negative count");
    SPC_REQUIRE(n_materials < 32768, SPC_ERR_INVALID, "spc_scene_upload: material ids are int16 (BDPTVertex.h:49)");
    uint64_t total = 0;
    for (int m = 0; m < n_meshes; m++) {
        const spc_mesh& me = meshes[m];
        SPC_REQUIRE(me.positions && me.indices, SPC_ERR_INVALID, "spc_scene_upload: mesh %d has null arrays", m);
        SPC_REQUIRE(me.light_id >= 0 ? me.light_id < n_lights : (me.material_id >= 0 && me.material_id < n_materials),
                    SPC_ERR_INVALID, "spc_scene_upload: mesh %d references material %d / light %d out of range", m,
                    me.material_id, me.light_id);
        total += me.n_triangles;
    }
    SPC_REQUIRE(total >= 1 && total < 0x7fffffffull, SPC_ERR_INVALID, "spc_scene_upload: %llu triangles", (unsigned long long)total);
    for (int i = 0; i < n_materials; i++)   // shade_pbr reads tex_desc[tex - 1]: a texture id above the table reads outside it
        SPC_REQUIRE(materials[i].base_color_tex.tex <= (uint64_t)n_textures, SPC_ERR_INVALID,
                    "spc_scene_upload: material %d names texture %llu of %d", i, (unsigned long long)materials[i].base_color_tex.tex, n_textures);
    const uint32_t n = (uint32_t)total;
    std::vector<float4> tri_pos((size_t)n * 3);
    std::vector<float2> tri_uv((size_t)n * 3);
    size_t p = 0;
    for (int m = 0; m < n_meshes; m++) {
        const spc_mesh& me = meshes[m];
        // per-mesh id words carried in the .w lanes: material (or -1), light id (or -1), mesh index
        const int mat = me.light_id >= 0 ? -1 : me.material_id;
        for (uint32_t t = 0; t < me.n_triangles; t++, p++) {
            for (int k = 0; k < 3; k++) {
                const uint32_t vi = me.indices[3 * (size_t)t + k];
                SPC_REQUIRE(vi < me.n_vertices, SPC_ERR_INVALID, "spc_scene_upload: mesh %d triangle %u index %u >= %u", m, t, vi, me.n_vertices);
                const float* q = me.positions + 3 * (size_t)vi;
                int wbits = k == 0 ? mat : (k == 1 ? me.light_id : m);
                float w;
                memcpy(&w, &wbits, 4);
                tri_pos[3 * p + k] = make_float4(q[0], q[1], q[2], w);
                tri_uv[3 * p + k] = me.texcoords ? make_float2(me.texcoords[2 * (size_t)vi], me.texcoords[2 * (size_t)vi + 1])
                                                 : make_float2(0.f, 0.f);
            }
        }
    }
    spc::SceneGeom& g = c.geom;
    g.n_prims = n;
    g.tri_pos.alloc(tri_pos.size());
    g.tri_uv.alloc(tri_uv.size());
    SPC_CUDA(cudaMemcpyAsync(g.tri_pos.p, tri_pos.data(), tri_pos.size() * sizeof(float4), cudaMemcpyHostToDevice, c.stream));
    SPC_CUDA(cudaMemcpyAsync(g.tri_uv.p, tri_uv.data(), tri_uv.size() * sizeof(float2), cudaMemcpyHostToDevice, c.stream));
    g.n_materials = n_materials;
    g.n_lights = n_lights;
    g.n_textures = n_textures;
    g.materials.alloc(n_materials);
    g.lights.alloc(n_lights);
    if (n_materials) SPC_CUDA(cudaMemcpyAsync(g.materials.p, materials, n_materials * sizeof(spc_pbr), cudaMemcpyHostToDevice, c.stream));
    if (n_lights) SPC_CUDA(cudaMemcpyAsync(g.lights.p, lights, n_lights * sizeof(spc_light), cudaMemcpyHostToDevice, c.stream));
    {
        std::vector<int4> desc(n_textures > 0 ? n_textures : 1);
        size_t bytes = 0;
        for (int t = 0; t < n_textures; t++) {
            SPC_REQUIRE(textures[t].rgba && textures[t].width > 0 && textures[t].height > 0, SPC_ERR_INVALID, "spc_scene_upload: texture %d is empty", t);
            desc[t] = make_int4((int)bytes, textures[t].width, textures[t].height, 0);
            bytes += (size_t)textures[t].width * textures[t].height * 4;
            SPC_REQUIRE(bytes < 0x7fffffffull, SPC_ERR_CAPACITY, "spc_scene_upload: textures exceed 2 GiB");
        }
        g.tex_data.alloc(bytes);
        g.tex_desc.alloc(desc.size());
        for (int t = 0; t < n_textures; t++)
            SPC_CUDA(cudaMemcpyAsync(g.tex_data.p + desc[t].x, textures[t].rgba, (size_t)textures[t].width * textures[t].height * 4,
                                     cudaMemcpyHostToDevice, c.stream));
        SPC_CUDA(cudaMemcpyAsync(g.tex_desc.p, desc.data(), desc.size() * sizeof(int4), cudaMemcpyHostToDevice, c.stream));
    }
    SPC_CUDA(cudaStreamSynchronize(c.stream));  // host staging vectors die at scope exit
    spc::build_material_tables(c);
    spc::build_bvh(c, g.tri_pos.p, n);
    c.has_scene = true;
    SPC_API_END
}

// Frame lanes on one GPU (and any other set of contexts on the same device) need ONE copy of the scene: the geometry, material, light
// and texture arrays and the BVH are read-only after the upload.  `ctx` becomes a second user of `owner`'s scene: no host staging, no
// copies, no second BVH build -- and one copy of the BVH in L2 instead of one per lane.  `owner` must outlive `ctx` (or `ctx` must
// upload / share another scene first) and must not upload a new scene while `ctx` renders.
int spc_scene_share(spc_context* ctx, spc_context* owner) {
    SPC_API_BEGIN
    SPC_REQUIRE(owner && owner != ctx, SPC_ERR_INVALID, "spc_scene_share: no owner context");
    const Context& o = owner->c;
    SPC_REQUIRE(o.has_scene, SPC_ERR_NO_SCENE, "spc_scene_share: the owner has no scene");
    SPC_REQUIRE(o.device == c.device, SPC_ERR_INVALID, "spc_scene_share: contexts live on devices %d and %d", o.device, c.device);
    SPC_CUDA(cudaStreamSynchronize(c.stream));   // nothing of this context may still read the buffers it is about to drop
    c.bvh.nodes.borrow(o.bvh.nodes);
    c.bvh.tris.borrow(o.bvh.tris);
    c.bvh.n_nodes = o.bvh.n_nodes;
    c.bvh.n_tris = o.bvh.n_tris;
    c.bvh_stats = o.bvh_stats;
    spc::SceneGeom& g = c.geom;
    const spc::SceneGeom& og = o.geom;
    g.tri_pos.borrow(og.tri_pos);
    g.tri_uv.borrow(og.tri_uv);
    g.materials.borrow(og.materials);
    g.mat_log_cc.borrow(og.mat_log_cc);
    g.lights.borrow(og.lights);
    g.tex_data.borrow(og.tex_data);
    g.tex_desc.borrow(og.tex_desc);
    g.n_prims = og.n_prims;
    g.n_materials = og.n_materials;
    g.n_lights = og.n_lights;
    g.n_textures = og.n_textures;
    for (int a = 0; a < 3; a++) { g.scene_lo[a] = og.scene_lo[a]; g.scene_hi[a] = og.scene_hi[a]; }
    c.has_scene = true;
    SPC_API_END
}

int spc_bvh_stats_get(spc_context* ctx, spc_bvh_stats* out) {
    SPC_API_BEGIN
    SPC_REQUIRE(out, SPC_ERR_INVALID, "spc_bvh_stats_get: out is null");
    SPC_REQUIRE(c.has_scene, SPC_ERR_NO_SCENE, "spc_bvh_stats_get: no scene uploaded");
    *out = c.bvh_stats;
    SPC_API_END
}

int spc_trace_batch_device(spc_context* ctx, const spc_ray* rays_dev, int64_t n, int ray_flags, spc_hit* hits_dev) {
    SPC_API_BEGIN
    SPC_REQUIRE(c.has_scene, SPC_ERR_NO_SCENE, "spc_trace_batch: no scene uploaded");
    SPC_REQUIRE(n >= 0 && (n == 0 || (rays_dev && hits_dev)), SPC_ERR_INVALID, "spc_trace_batch: bad arguments");
    spc::launch_trace_closest(c, rays_dev, n, ray_flags, hits_dev, nullptr);
    SPC_API_END
}

int spc_trace_batch(spc_context* ctx, const spc_ray* rays_host, int64_t n, int ray_flags, spc_hit* hits_host) {
    SPC_API_BEGIN
    SPC_REQUIRE(c.has_scene, SPC_ERR_NO_SCENE, "spc_trace_batch: no scene uploaded");
    SPC_REQUIRE(n >= 0 && (n == 0 || (rays_host && hits_host)), SPC_ERR_INVALID, "spc_trace_batch: bad arguments");
    if (n > 0) {
        c.scratch_rays.alloc(n);
        c.scratch_hits.alloc(n);
        pipelined_host_batch(
            c, rays_host, n, [&](int64_t off, int64_t m) { spc::launch_trace_closest(c, c.scratch_rays.p + off, m, ray_flags, c.scratch_hits.p + off, nullptr); },
            [&](int64_t off, int64_t m, cudaStream_t st) {
                SPC_CUDA(cudaMemcpyAsync(hits_host + off, c.scratch_hits.p + off, m * sizeof(spc_hit), cudaMemcpyDeviceToHost, st));
            });
    }
    SPC_API_END
}

int spc_occlusion_batch_device(spc_context* ctx, const spc_ray* rays_dev, int64_t n, uint8_t* visible_dev) {
    SPC_API_BEGIN
    SPC_REQUIRE(c.has_scene, SPC_ERR_NO_SCENE, "spc_occlusion_batch: no scene uploaded");
    SPC_REQUIRE(n >= 0 && (n == 0 || (rays_dev && visible_dev)), SPC_ERR_INVALID, "spc_occlusion_batch: bad arguments");
    spc::launch_trace_occlusion(c, rays_dev, n, visible_dev, nullptr);
    SPC_API_END
}

int spc_occlusion_batch(spc_context* ctx, const spc_ray* rays_host, int64_t n, uint8_t* visible_host) {
    SPC_API_BEGIN
    SPC_REQUIRE(c.has_scene, SPC_ERR_NO_SCENE, "spc_occlusion_batch: no scene uploaded");
    SPC_REQUIRE(n >= 0 && (n == 0 || (rays_host && visible_host)), SPC_ERR_INVALID, "spc_occlusion_batch: bad arguments");
    if (n > 0) {
        c.scratch_rays.alloc(n);
        c.scratch_vis.alloc(n);
        pipelined_host_batch(
            c, rays_host, n, [&](int64_t off, int64_t m) { spc::launch_trace_occlusion(c, c.scratch_rays.p + off, m, c.scratch_vis.p + off, nullptr); },
            [&](int64_t off, int64_t m, cudaStream_t st) { SPC_CUDA(cudaMemcpyAsync(visible_host + off, c.scratch_vis.p + off, m, cudaMemcpyDeviceToHost, st)); });
    }
    SPC_API_END
}

static int counted_finish(Context& c, int64_t n, spc_trace_counters* out) {
    unsigned long long h[2];
    SPC_CUDA(cudaMemcpyAsync(h, c.counters.p, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    out->rays = (uint64_t)n;
    out->nodes_visited = h[0];
    out->tris_tested = h[1];
    return 0;
}

int spc_trace_batch_counted(spc_context* ctx, const spc_ray* rays_dev, int64_t n, int ray_flags, spc_hit* hits_dev,
                            spc_trace_counters* out_host) {
    SPC_API_BEGIN
    SPC_REQUIRE(c.has_scene, SPC_ERR_NO_SCENE, "spc_trace_batch_counted: no scene uploaded");
    SPC_REQUIRE(n >= 0 && out_host && (n == 0 || (rays_dev && hits_dev)), SPC_ERR_INVALID, "spc_trace_batch_counted: bad arguments");
    SPC_CUDA(cudaMemsetAsync(c.counters.p, 0, 16, c.stream));
    spc::launch_trace_closest(c, rays_dev, n, ray_flags, hits_dev, c.counters.p);
    counted_finish(c, n, out_host);
    SPC_API_END
}

int spc_occlusion_batch_counted(spc_context* ctx, const spc_ray* rays_dev, int64_t n, uint8_t* visible_dev,
                                spc_trace_counters* out_host) {
    SPC_API_BEGIN
    SPC_REQUIRE(c.has_scene, SPC_ERR_NO_SCENE, "spc_occlusion_batch_counted: no scene uploaded");
    SPC_REQUIRE(n >= 0 && out_host && (n == 0 || (rays_dev && visible_dev)), SPC_ERR_INVALID, "spc_occlusion_batch_counted: bad arguments");
    SPC_CUDA(cudaMemsetAsync(c.counters.p, 0, 16, c.stream));
    spc::launch_trace_occlusion(c, rays_dev, n, visible_dev, c.counters.p);
    counted_finish(c, n, out_host);
    SPC_API_END
}

}  // extern "C"
