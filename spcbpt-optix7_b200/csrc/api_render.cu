// api_render.cu -- extern "C" entry points of the render path (launch seam + MyThrustOp seam).
#include <cstring>
#include "common.cuh"

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

extern "C" {

int spc_set_params(spc_context* ctx, const spc_params* params) {
    SPC_API_BEGIN
    SPC_REQUIRE(params, SPC_ERR_INVALID, "spc_set_params: params is null");
    SPC_REQUIRE(!params->sky.valid, SPC_ERR_INVALID, "spc_set_params: environment lighting (sky.valid) is not supported (unfinished in the reference, readme.md:28)");
    c.params = *params;
    c.has_params = true;
    SPC_API_END
}

int spc_launch(spc_context* ctx, int kind, int width, int height) {
    SPC_API_BEGIN
    SPC_REQUIRE(c.has_scene, SPC_ERR_NO_SCENE, "spc_launch: no scene uploaded");
    switch (kind) {
        case SPC_LAUNCH_LIGHT_TRACE:
            SPC_REQUIRE(c.has_params && width == c.params.lt.num_core && height == 1, SPC_ERR_INVALID,
                        "spc_launch(light trace): launch size must be (lt.num_core, 1)");
            spc::launch_light_trace(c);
            break;
        case SPC_LAUNCH_PRETRACE:
            SPC_REQUIRE(c.has_params && width == c.params.pre_tracer.num_core && height == 1, SPC_ERR_INVALID,
                        "spc_launch(pretrace): launch size must be (pre_tracer.num_core, 1)");
            spc::launch_pretrace(c);
            break;
        case SPC_LAUNCH_PT:
            spc::launch_pt(c, width, height);
            break;
        case SPC_LAUNCH_SPCBPT_EYE:
            spc::launch_eye_pass(c, width, height);
            break;
        default:
            SPC_REQUIRE(false, SPC_ERR_INVALID, "spc_launch: raygen kind %d is not available in this build", kind);
    }
    SPC_API_END
}

int spc_launch_named(spc_context* ctx, const char* name, int width, int height) {
    if (!name) {
        spc::set_error("spc_launch_named: null name");
        return SPC_ERR_INVALID;
    }
    int kind = -1;
    if (!strcmp(name, "pt")) kind = SPC_LAUNCH_PT;
    else if (!strcmp(name, "SPCBPT_eye")) kind = SPC_LAUNCH_SPCBPT_EYE;
    else if (!strcmp(name, "light trace")) kind = SPC_LAUNCH_LIGHT_TRACE;
    else if (!strcmp(name, "pretrace")) kind = SPC_LAUNCH_PRETRACE;
    if (kind < 0) {
        spc::set_error("spc_launch_named: unknown raygen '%s'", name);   // the reference throws (Scene.cpp:1693-1696)
        return SPC_ERR_INVALID;
    }
    return spc_launch(ctx, kind, width, height);
}

int spc_set_seed_offset(spc_context* ctx, uint32_t offset) {
    SPC_API_BEGIN
    c.seed_offset = offset;
    SPC_API_END
}

int spc_set_seed_mapping(spc_context* ctx, uint32_t offset, uint32_t stride) {
    SPC_API_BEGIN
    SPC_REQUIRE(stride >= 1, SPC_ERR_INVALID, "spc_set_seed_mapping: stride must be >= 1");
    c.seed_offset = offset;
    c.seed_stride = stride;
    SPC_API_END
}

int spc_set_tile_partition(spc_context* ctx, int gpu_idx, int num_gpus) {
    SPC_API_BEGIN
    SPC_REQUIRE(num_gpus >= 1 && gpu_idx >= 0 && gpu_idx < num_gpus, SPC_ERR_INVALID, "spc_set_tile_partition: gpu %d of %d", gpu_idx, num_gpus);
    c.tile_rank = gpu_idx;
    c.tile_world = num_gpus;
    SPC_API_END
}

int spc_set_trace_blocks(spc_context* ctx, int blocks_per_sm) {
    SPC_API_BEGIN
    SPC_REQUIRE(blocks_per_sm >= 0 && blocks_per_sm <= 32, SPC_ERR_INVALID, "spc_set_trace_blocks: %d is outside 0..32", blocks_per_sm);
    c.trace_blocks_per_sm = blocks_per_sm;
    SPC_API_END
}

int spc_merge_accum(spc_context* ctx, const spc_float4* const* accum_dev, const float* weights, int n, int n_pixels, spc_float4* out_accum_dev,
                    uint32_t* out_frame_dev) {
    SPC_API_BEGIN
    spc::merge_accum(c, accum_dev, weights, n, n_pixels, out_accum_dev, out_frame_dev);
    SPC_API_END
}

// name -> field table of the context switches
static int64_t* option_slot(Context& c, const char* name, int64_t* lo, int64_t* hi) {
    struct Opt { const char* name; int64_t* p; int64_t lo, hi; };
    const Opt table[] = {
        {"reference_search", &c.opt[spc::OPT_REFERENCE_SEARCH], 0, 1},
        {"blocking_sync", &c.opt[spc::OPT_BLOCKING_SYNC], 0, 1},
        {"count_canonical", &c.opt[spc::OPT_COUNT_CANONICAL], 0, 1},
        {"stage_timing", &c.opt[spc::OPT_STAGE_TIMING], 0, 1},
        {"light_trace_mode", &c.opt[spc::OPT_LIGHT_TRACE_MODE], 0, 1},
        {"tail_threshold", &c.opt[spc::OPT_TAIL_THRESHOLD], -1, 1 << 30},
        {"sort_hits", &c.opt[spc::OPT_SORT_HITS], 0, 1},
        {"train_reserve_paths", &c.opt[spc::OPT_TRAIN_RESERVE], 0, 1 << 30},
    };
    for (const Opt& o : table)
        if (!strcmp(o.name, name)) {
            *lo = o.lo; *hi = o.hi;
            return o.p;
        }
    return nullptr;
}

int spc_set_option(spc_context* ctx, const char* name, int64_t value) {
    SPC_API_BEGIN
    SPC_REQUIRE(name, SPC_ERR_INVALID, "spc_set_option: null name");
    int64_t lo, hi;
    int64_t* p = option_slot(c, name, &lo, &hi);
    SPC_REQUIRE(p, SPC_ERR_INVALID, "spc_set_option: unknown option '%s'", name);
    SPC_REQUIRE(value >= lo && value <= hi, SPC_ERR_INVALID, "spc_set_option: %s = %lld outside [%lld, %lld]", name, (long long)value, (long long)lo, (long long)hi);
    *p = value;
    SPC_API_END
}

int spc_get_option(spc_context* ctx, const char* name, int64_t* value) {
    SPC_API_BEGIN
    SPC_REQUIRE(name && value, SPC_ERR_INVALID, "spc_get_option: null argument");
    int64_t lo, hi;
    int64_t* p = option_slot(c, name, &lo, &hi);
    SPC_REQUIRE(p, SPC_ERR_INVALID, "spc_get_option: unknown option '%s'", name);
    *value = *p;
    SPC_API_END
}

int spc_eye_stats_get(spc_context* ctx, spc_eye_stats* out) {
    SPC_API_BEGIN
    SPC_REQUIRE(out, SPC_ERR_INVALID, "spc_eye_stats_get: out is null");
    spc::eye_stats(c, out);
    SPC_API_END
}

int spc_set_debug_outputs(spc_context* ctx, int32_t* first_prim_dev, int32_t* first_label_dev) {
    SPC_API_BEGIN
    c.dbg_first_prim = first_prim_dev;
    c.dbg_first_label = first_label_dev;
    SPC_API_END
}

int spc_lvc_process(spc_context* ctx, const spc_vertex* lvc_dev, const uint8_t* valid_dev, int count_range, spc_subspace_sampler* out_host) {
    SPC_API_BEGIN
    spc::lvc_process(c, lvc_dev, valid_dev, count_range, out_host);
    SPC_API_END
}

}  // extern "C"
