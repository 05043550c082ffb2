// layout_check.cu -- compile-time proof that the POD mirrors in include/spcbpt_b200.h keep the
// byte layout of the reference's host<->device structs (sizes/offsets probed from the reference
// headers, SURVEY.md section 8 intro; re-checked against the headers themselves by
// oracle/ref_shim/ref_host.cpp -> tests/test_layout.py).
#include <cstddef>
#include "../../include/spcbpt_b200.h"

static_assert(sizeof(spc_texture_ref) == 40, "MaterialData::Texture");
static_assert(sizeof(spc_pbr) == 144 && alignof(spc_pbr) == 16, "MaterialData::Pbr");
static_assert(offsetof(spc_pbr, base_color_tex) == 56 && offsetof(spc_pbr, metallic_roughness_tex) == 96 &&
                  offsetof(spc_pbr, brdf) == 136, "MaterialData::Pbr fields");
static_assert(sizeof(spc_light) == 80, "Light");
static_assert(offsetof(spc_light, corner) == 16 && offsetof(spc_light, emission) == 52 &&
                  offsetof(spc_light, normal) == 64 && offsetof(spc_light, area) == 76, "Light::quad");
static_assert(sizeof(spc_vertex) == 120, "BDPTVertex");
static_assert(offsetof(spc_vertex, RMIS_pointer_3) == 60 && offsetof(spc_vertex, uv) == 72 &&
                  offsetof(spc_vertex, RMIS_pointer) == 80 && offsetof(spc_vertex, pdf) == 92 &&
                  offsetof(spc_vertex, materialId) == 104 && offsetof(spc_vertex, subspaceId) == 106 &&
                  offsetof(spc_vertex, type) == 112 && offsetof(spc_vertex, isOrigin) == 114 &&
                  offsetof(spc_vertex, isLastVertex_direction) == 118, "BDPTVertex fields");
static_assert(sizeof(spc_tree_node) == 56 && offsetof(spc_tree_node, label) == 44 && offsetof(spc_tree_node, leaf) == 52, "tree_node");
static_assert(sizeof(spc_divide_weight) == 40, "divide_weight");
static_assert(sizeof(spc_subspace) == 20, "Subspace");
static_assert(sizeof(spc_buffer_view) == 16, "BufferView");
static_assert(sizeof(spc_light_trace_params) == 40, "LightTraceParams");
static_assert(sizeof(spc_pretrace_params) == 32, "PreTraceParams");
static_assert(sizeof(spc_subspace_sampler) == 40, "SubspaceSampler");
static_assert(sizeof(spc_subspace_macro_info) == 40, "subspaceMacroInfo");
static_assert(sizeof(spc_env_info) == 56, "envInfo");
static_assert(sizeof(spc_params) == 352, "MyParams");
static_assert(offsetof(spc_params, width) == 0 && offsetof(spc_params, height) == 4 &&
                  offsetof(spc_params, subframe_index) == 8 && offsetof(spc_params, accum_buffer) == 16 &&
                  offsetof(spc_params, frame_buffer) == 24 && offsetof(spc_params, max_depth) == 32 &&
                  offsetof(spc_params, eye) == 36 && offsetof(spc_params, U) == 48 && offsetof(spc_params, V) == 60 &&
                  offsetof(spc_params, W) == 72 && offsetof(spc_params, lights) == 88 &&
                  offsetof(spc_params, materials) == 104 && offsetof(spc_params, miss_color) == 120 &&
                  offsetof(spc_params, handle) == 136 && offsetof(spc_params, lt) == 144 &&
                  offsetof(spc_params, sampler) == 184 && offsetof(spc_params, pre_tracer) == 224 &&
                  offsetof(spc_params, subspace_info) == 256 && offsetof(spc_params, sky) == 296, "MyParams fields");
static_assert(sizeof(spc_ray) == 32 && sizeof(spc_hit) == 16, "ray batch records");
static_assert(sizeof(spc_train_path) == 48 && offsetof(spc_train_path, sample_pdf) == 12 && offsetof(spc_train_path, begin_ind) == 20 &&
                  offsetof(spc_train_path, pixel_x) == 32 && offsetof(spc_train_path, valid) == 40, "TrainData::pathInfo_sample");
static_assert(sizeof(spc_train_conn) == 92 && offsetof(spc_train_conn, peak_pdf) == 72 && offsetof(spc_train_conn, path_id) == 76 &&
                  offsetof(spc_train_conn, label_A) == 80 && offsetof(spc_train_conn, label_B) == 84 && offsetof(spc_train_conn, valid) == 88 &&
                  offsetof(spc_train_conn, light_source) == 89, "TrainData::pathInfo_node");
