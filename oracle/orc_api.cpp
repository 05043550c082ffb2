// orc_api.cpp -- C entry points of the CPU oracle (loaded with ctypes by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs ONLY).
// TEST INFRASTRUCTURE: the product library never links or calls this.
#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>
#include "orc_scene.h"

using namespace orc;

static void parallel_for(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& body) {
    if (threads <= 1 || n < 1024) {
        body(0, n);
        return;
    }
    std::atomic<int64_t> next(0);
    const int64_t chunk = 4096;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&]() {
            for (;;) {
                const int64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                body(b, std::min(n, b + chunk));
            }
        });
    for (auto& th : pool) th.join();
}

extern "C" {

#define ORC_API __attribute__((visibility("default")))

ORC_API uint32_t orc_tea(uint32_t rounds, uint32_t v0, uint32_t v1) { return tea(rounds, v0, v1); }

// draws n floats from the LCG stream starting at *state; updates *state
ORC_API void orc_rnd_stream(uint32_t* state, int n, float* out) {
    for (int i = 0; i < n; i++) out[i] = rnd(*state);
}

ORC_API void* orc_scene_create(const spc_mesh* meshes, int n_meshes, const spc_pbr* materials, int n_materials,
                               const spc_light* lights, int n_lights, const spc_texture* textures, int n_textures) {
    Scene* s = new Scene();
    for (int m = 0; m < n_meshes; m++) {
        const spc_mesh& me = meshes[m];
        for (uint32_t t = 0; t < me.n_triangles; t++) {
            Tri tr;
            f3 p[3];
            for (int k = 0; k < 3; k++) {
                const uint32_t vi = me.indices[3 * (size_t)t + k];
                p[k] = mk3(me.positions[3 * (size_t)vi], me.positions[3 * (size_t)vi + 1], me.positions[3 * (size_t)vi + 2]);
                tr.uv[k][0] = me.texcoords ? me.texcoords[2 * (size_t)vi] : 0.f;
                tr.uv[k][1] = me.texcoords ? me.texcoords[2 * (size_t)vi + 1] : 0.f;
            }
            tr.v0 = p[0]; tr.v1 = p[1]; tr.v2 = p[2];
            tr.e1 = p[1] - p[0];
            tr.e2 = p[2] - p[0];
            tr.material = me.light_id >= 0 ? -1 : me.material_id;
            tr.light = me.light_id;
            tr.mesh = m;
            s->tris.push_back(tr);
        }
    }
    s->materials.assign(materials, materials + n_materials);
    s->lights.assign(lights, lights + n_lights);
    for (int t = 0; t < n_textures; t++) {
        Texture tx;
        tx.w = textures[t].width;
        tx.h = textures[t].height;
        tx.rgba.assign(textures[t].rgba, textures[t].rgba + (size_t)tx.w * tx.h * 4);
        s->textures.push_back(std::move(tx));
    }
    s->build_bvh();
    return s;
}

ORC_API void orc_scene_destroy(void* sc) { delete (Scene*)sc; }
ORC_API int orc_scene_num_prims(void* sc) { return (int)((Scene*)sc)->tris.size(); }

ORC_API void orc_trace_batch(void* sc, const spc_ray* rays, int64_t n, int ray_flags, spc_hit* hits, int brute, int threads) {
    const Scene* s = (const Scene*)sc;
    const bool cull = (ray_flags & SPC_RAYFLAG_CULL_BACK_FACING) != 0;
    parallel_for(n, threads, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; i++) {
            Hit h;
            s->closest(mk3(rays[i].ox, rays[i].oy, rays[i].oz), mk3(rays[i].dx, rays[i].dy, rays[i].dz), rays[i].tmin,
                       rays[i].tmax, cull, h, brute != 0);
            hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v; hits[i].prim = h.prim;
        }
    });
}

ORC_API void orc_occlusion_batch(void* sc, const spc_ray* rays, int64_t n, uint8_t* visible, int brute, int threads) {
    const Scene* s = (const Scene*)sc;
    parallel_for(n, threads, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; i++)
            visible[i] = s->occluded(mk3(rays[i].ox, rays[i].oy, rays[i].oz), mk3(rays[i].dx, rays[i].dy, rays[i].dz),
                                     rays[i].tmin, rays[i].tmax, brute != 0) ? 0 : 1;
    });
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// render path (orc_render.cpp)
// ---------------------------------------------------------------------------------------------
#include "orc_render.h"

static Frame make_frame(void* sc, const spc_params* p, int K, int connections, int max_depth) {
    Frame fr;
    fr.sc = (const Scene*)sc;
    fr.p = *p;
    fr.K = K;
    fr.connections = connections;
    fr.max_depth = max_depth > 0 ? max_depth : 50;
    return fr;
}

extern "C" {

ORC_API void orc_set_jitter_rtl(int v) { g_jitter_rtl = v; }

// optixLaunch of "light trace" (optixPathTracer.cpp:491-514): one sequential core per launch index
// (connections = CONNECTION_N, optixPathTracer.h:33: the light sub-path's MIS recursion multiplies by it, cuProg.h:70-78)
ORC_API void orc_light_trace_c(void* sc, const spc_params* p, int K, int connections, int max_depth, int threads);
ORC_API void orc_light_trace(void* sc, const spc_params* p, int K, int max_depth, int threads) { orc_light_trace_c(sc, p, K, 3, max_depth, threads); }
ORC_API void orc_light_trace_c(void* sc, const spc_params* p, int K, int connections, int max_depth, int threads) {
    const Frame fr = make_frame(sc, p, K, connections, max_depth);
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            const int c = next.fetch_add(1);
            if (c >= fr.p.lt.num_core) break;
            light_trace_core(fr, c);
        }
    };
    if (threads <= 1) worker();
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++) pool.emplace_back(worker);
        for (auto& th : pool) th.join();
    }
}

// MyThrustOp::LVC_Process (device_thrust.cu:241-332)
ORC_API void orc_lvc_process(const spc_vertex* lvc, const uint8_t* valid, int n, int K, spc_subspace* subspace, float* cmfs,
                             int* jump, int* vertex_count, int* path_count) {
    lvc_process(lvc, valid, n, K, subspace, cmfs, jump, vertex_count, path_count);
}

// optixLaunch of "SPCBPT_eye" (optixPathTracer.cpp:609-635); optional per-pixel bounce-0 prim / subspace ids
ORC_API void orc_eye_pass(void* sc, const spc_params* p, int K, int connections, int max_depth, int threads, int* first_prim,
                          int* first_label) {
    const Frame fr = make_frame(sc, p, K, connections, max_depth);
    const int W = (int)p->width, H = (int)p->height;
    parallel_for((int64_t)W * H, threads, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; i++)
            eye_pixel(fr, (int)(i % W), (int)(i / W), first_prim ? first_prim + i : nullptr, first_label ? first_label + i : nullptr);
    });
}

// optixLaunch of "pt" (the comparison integrator, raygen.cu:71-170)
ORC_API void orc_pt_pass(void* sc, const spc_params* p, int K, int threads) {
    const Frame fr = make_frame(sc, p, K, 3, 0);
    const int W = (int)p->width, H = (int)p->height;
    parallel_for((int64_t)W * H, threads, [&](int64_t b, int64_t e) {
        for (int64_t i = b; i < e; i++) pt_pixel(fr, (int)(i % W), (int)(i / W));
    });
}

// stage-wise entry points -----------------------------------------------------------------------
ORC_API void orc_bsdf(void* sc, int material_id, const float* color3, const float* N, const float* V, const float* L, uint32_t* seed,
                      float* eval3, float* pdf1, float* sample3) {
    Pbr m = load_pbr(*(const Scene*)sc, material_id);
    if (color3) m.base_color = mk3(color3[0], color3[1], color3[2]);
    const f3 n = mk3(N[0], N[1], N[2]), v = mk3(V[0], V[1], V[2]), l = mk3(L[0], L[1], L[2]);
    const f3 e = bsdf_eval(m, n, v, l);
    eval3[0] = e.x; eval3[1] = e.y; eval3[2] = e.z;
    *pdf1 = bsdf_pdf(m, n, v, l);
    const f3 s = bsdf_sample(m, n, v, *seed);
    sample3[0] = s.x; sample3[1] = s.y; sample3[2] = s.z;
}

ORC_API void orc_classify(const spc_tree_node* tree, const float* pos, const float* nrm, int n, int* labels) {
    for (int i = 0; i < n; i++)
        labels[i] = tree_label(tree, mk3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), mk3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
}

// connectVertex_SPCBPT (raygen.cu:253-303) on caller-provided vertex pairs: out3 = contribution, w = MIS weight
ORC_API void orc_connect(void* sc, const spc_params* p, int K, int connections, const spc_vertex* eye, const spc_vertex* light, int n,
                         float* out3, float* w) {
    const Frame fr = make_frame(sc, p, K, connections, 0);
    for (int i = 0; i < n; i++) {
        f3 c;
        w[i] = connect_mis_and_eval(fr, eye[i], light[i], c);
        out3[3 * i] = c.x; out3[3 * i + 1] = c.y; out3[3 * i + 2] = c.z;
    }
}

}  // extern "C"


// ---------------------------------------------------------------------------------------------
// training path (orc_train.cpp)
// ---------------------------------------------------------------------------------------------
#include "orc_train.h"

extern "C" {

// optixLaunch of "pretrace" (optixPathTracer.cpp:523-545): one training path per launch index
ORC_API void orc_pretrace(void* sc, const spc_params* p, int K, int max_depth, int threads) {
    const Frame fr = make_frame(sc, p, K, 3, max_depth);
    parallel_for(fr.p.pre_tracer.num_core, threads > 1 && fr.p.pre_tracer.num_core >= 1024 ? threads : 1,
                 [&](int64_t b, int64_t e) { for (int64_t i = b; i < e; i++) pretrace_core(fr, (int)i); });
}

ORC_API void* orc_ts_create() { return new TrainSet(); }
ORC_API void orc_ts_destroy(void* ts) { delete (TrainSet*)ts; }
ORC_API int orc_ts_gather(void* ts, const spc_train_path* paths, int n_paths, const spc_train_conn* conns, int n_conns) {
    return ((TrainSet*)ts)->gather(paths, n_paths, conns, n_conns);
}
ORC_API int orc_ts_sizes(void* ts, int* n_paths, int* n_conns) {
    *n_paths = (int)((TrainSet*)ts)->paths.size();
    *n_conns = (int)((TrainSet*)ts)->conns.size();
    return 0;
}
ORC_API void orc_ts_read(void* ts, spc_train_path* paths, spc_train_conn* conns) {
    TrainSet* t = (TrainSet*)ts;
    if (paths) memcpy(paths, t->paths.data(), t->paths.size() * sizeof(spc_train_path));
    if (conns) memcpy(conns, t->conns.data(), t->conns.size() * sizeof(spc_train_conn));
}
ORC_API void orc_ts_reweight(void* ts) { ((TrainSet*)ts)->reweight(); }
ORC_API int orc_ts_tree_points(void* ts, int eye_side, int max_size, spc_divide_weight* out, int cap) {
    const std::vector<spc_divide_weight> v = ((TrainSet*)ts)->tree_points(eye_side != 0, max_size);
    if ((int)v.size() <= cap) memcpy(out, v.data(), v.size() * sizeof(spc_divide_weight));
    return (int)v.size();
}
ORC_API void orc_ts_label(void* ts, const spc_tree_node* eye_tree, const spc_tree_node* light_tree) { ((TrainSet*)ts)->label(eye_tree, light_tree); }
ORC_API void* orc_q_create(int K) { return new QEstimator(K); }
ORC_API void orc_q_destroy(void* q) { delete (QEstimator*)q; }
ORC_API int orc_q_add(void* q, const spc_vertex* lvc, const uint8_t* valid, int n) { return ((QEstimator*)q)->add(lvc, valid, n); }
ORC_API void orc_q_zero_handle(void* q) { ((QEstimator*)q)->zero_handle(); }
ORC_API void orc_q_read(void* q, float* out) { memcpy(out, ((QEstimator*)q)->Q.data(), ((QEstimator*)q)->K * sizeof(float)); }

struct OrcTrainData { TrainData td; };
ORC_API void* orc_ts_build_train_data(void* ts, int n_samples, const float* Q, int K) {
    OrcTrainData* d = new OrcTrainData();
    ((TrainSet*)ts)->build_train_data(n_samples, Q, K, d->td);
    return d;
}
ORC_API void orc_td_destroy(void* td) { delete (OrcTrainData*)td; }
ORC_API void orc_td_sizes(void* td, int* N, int* M, float* threshold) {
    *N = ((OrcTrainData*)td)->td.N; *M = ((OrcTrainData*)td)->td.M; *threshold = ((OrcTrainData*)td)->td.outlier_threshold;
}
ORC_API void orc_td_read(void* td, float* f_square, float* pdf0, int* P2N, float* peak, int* label_E, int* label_P) {
    const TrainData& t = ((OrcTrainData*)td)->td;
    memcpy(f_square, t.f_square.data(), t.N * 4); memcpy(pdf0, t.pdf0.data(), t.N * 4); memcpy(P2N, t.P2N.data(), t.N * 4);
    memcpy(peak, t.peak.data(), t.M * 4); memcpy(label_E, t.label_E.data(), t.M * 4); memcpy(label_P, t.label_P.data(), t.M * 4);
}
ORC_API void orc_ts_gamma_histogram(void* ts, int K, float* G) {
    std::vector<float> g;
    ((TrainSet*)ts)->gamma_histogram(K, g);
    memcpy(G, g.data(), g.size() * 4);
}
ORC_API void orc_train_gamma(void* td, int K, float* G, int batch_size, int epochs, float lr, float* loss_log, int loss_cap, int* n_loss) {
    std::vector<float> g(G, G + (size_t)K * K), loss;
    train_gamma(((OrcTrainData*)td)->td, K, g, batch_size, epochs, lr, 0.2, &loss);
    memcpy(G, g.data(), g.size() * 4);
    if (n_loss) *n_loss = (int)loss.size();
    if (loss_log) memcpy(loss_log, loss.data(), std::min((size_t)loss_cap, loss.size()) * 4);
}
ORC_API void orc_gamma_to_cmf(const float* G, int K, float* cmf) {
    std::vector<float> g(G, G + (size_t)K * K), c;
    gamma_to_cmf(g, K, 0.2f, c);
    memcpy(cmf, c.data(), c.size() * 4);
}

}  // extern "C"
