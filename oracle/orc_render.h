// orc_render.h -- CPU oracle: scalar restatement of the reference's SPCBPT render path.
// TEST INFRASTRUCTURE ONLY.  Every function cites the reference lines it restates (paths under
// /root/reference/src/OptiXPathTracer unless noted).  Pinned bit-for-bit against the reference's
// own sources compiled for the host (oracle/_ref/libref_host.so) by tests/test_oracle_vs_ref.py.
//
// Differences from the reference, all deliberate and documented in DESIGN.md:
//   * NUM_SUBSPACE / CONNECTION_N / max depth are runtime values (K, connections, max_depth);
//   * ray queries use the intersection contract of orc_scene.cpp (OptiX is closed source);
//   * textures use an fp32 bilinear/wrap fetch (CUDA's 9-bit filter weights are not reproducible);
//   * vertices are zero-initialised where the reference leaves stack garbage in unused fields.
#pragma once
#include "orc_scene.h"

namespace orc {

struct Pbr {   // the fields of MaterialData::Pbr the BSDF reads (src/cuda/MaterialData.h:78-97)
    f3 base_color;
    float metallic, roughness, specular, specularTint, subsurface, sheen, sheenTint, clearcoat, clearcoatGloss;
    bool brdf;
};

struct Frame {   // what the programs read from MyParams (optixPathTracer.h:191-199)
    const Scene* sc;
    spc_params p;          // host pointers
    int K;                 // NUM_SUBSPACE
    int connections;       // CONNECTION_N
    int max_depth;         // literal 50 in raygen.cu:361,668
};

extern int g_jitter_rtl;   // 1: evaluate the two jitter draws right-to-left like g++ (pinning vs libref_host only)

Pbr load_pbr(const Scene& sc, int material_id);
f3 bsdf_eval(const Pbr& m, f3 N, f3 V, f3 L);
f3 bsdf_sample(const Pbr& m, f3 N, f3 V, uint32_t& seed);
float bsdf_pdf(const Pbr& m, f3 N, f3 V, f3 L);
int tree_label(const spc_tree_node* root, f3 position, f3 normal);
float connect_mis_and_eval(const Frame& fr, const spc_vertex& a, const spc_vertex& b, f3& out);

void light_trace_core(const Frame& fr, int core);
void eye_pixel(const Frame& fr, int px, int py, int* first_prim, int* first_label);
void pt_pixel(const Frame& fr, int px, int py);   // the "pt" comparison integrator
void lvc_process(const spc_vertex* lvc, const uint8_t* valid, int n, int K, spc_subspace* subspace, float* cmfs, int* jump,
                 int* vertex_count, int* path_count);

}  // namespace orc
