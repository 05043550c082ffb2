// orc_internal.h -- declarations shared by the oracle's translation units (orc_render.cpp, orc_train.cpp).
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include "orc_render.h"

namespace orc {

struct LightSample {
    f3 position, emission, direction;
    float uvx, uvy, pdf, dir_pdf;
    int subspaceId;
    const spc_light* light;
};

// path state: BDPTPath keeps 3 vertices in a ring (BDPTVertex.h:72-117); the programs only ever touch
// current/last/next, and `next` carries the two values pre-loaded by the previous hit (flux = BSDF value,
// singlePdf), hit_program.cu:286-287 + :335
struct Path {
    spc_vertex v[3];
    int size;
    spc_vertex& cur() { return v[(size - 1) % 3]; }
    spc_vertex& next() { return v[size % 3]; }
    spc_vertex& last() { return v[(size - 2) % 3]; }
};
struct Payload {                                     // Tracer::PayloadBDPTVertex, cuProg.h:303-323
    Path path;
    f3 origin, ray_direction;
    float pdf;
    uint32_t seed;
    int depth;
    bool done;
    void clear() { path.size = 0; depth = 0; done = false; }
};


void light_reverse_sample(const Frame& fr, const spc_light& L, float r1, float r2, LightSample& s);
void light_sample_pos(const Frame& fr, const spc_light& L, uint32_t& seed, LightSample& s);
int pick_light(const Frame& fr, uint32_t& seed);
void init_vertex_from_light_sample(const LightSample& s, spc_vertex& v);
Pbr vertex_mat(const Frame& fr, const spc_vertex& v);
void trace_subpath(const Frame& fr, Payload& prd, f3 o, f3 d, bool light_side);
bool visibility_test(const Frame& fr, f3 pos_A, f3 pos_B);
bool invalid3(f3 a);

static inline f3 ld(const spc_float3& v) { return f3{v.x, v.y, v.z}; }
static inline void st(spc_float3& d, f3 v) { d.x = v.x; d.y = v.y; d.z = v.z; }

}  // namespace orc
