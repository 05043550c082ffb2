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
