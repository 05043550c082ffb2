// pt.cu -- the reference's comparison integrator "pt" (Space toggles between it and SPCBPT_eye, operation.md:5):
// unidirectional path tracing with next-event estimation and MIS.  Replaces optixLaunch of __raygen__pinhole
// (raygen.cu:71-170) with __closesthit__radiance (hit_program.cu:439-552), __closesthit__lightsource (:148-180) and
// __miss__constant_radiance (raygen.cu:687-696).  Also the in-engine ground-truth generator for relMSE checks.
// One lane per pixel with the traversal inlined (closest hit + one shadow ray per bounce).
#include "shade.cuh"
#include "traverse.cuh"

namespace spc {

DevFrame make_dev_frame(Context& c);

__device__ __forceinline__ unsigned pt_quantize8(float x) {
    x = clampf(x, 0.0f, 1.0f);
    return min((unsigned)(x * 256.0f), 255u);
}
__device__ __forceinline__ float pt_to_srgb(float c) {
    const float invGamma = 1.0f / 2.4f;
    const float powed = cm_powf(c, invGamma);
    return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f;
}

constexpr int kPtPixBlock = 64;

__global__ void __launch_bounds__(kPtPixBlock) k_pt(const DevFrame fr, int n_pix) {
    __shared__ uint2 s_stack[kSmStack * kPtPixBlock];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);   // before any thread leaves
    const int i = blockIdx.x * kPtPixBlock + threadIdx.x;
    if (i >= n_pix) return;
    uint2* stack = s_stack + threadIdx.x;
    unsigned cn = 0, ct = 0;
    const unsigned W = fr.p.width, H = fr.p.height;
    const unsigned x = (unsigned)i % W, y = (unsigned)i / W;
    const uint32_t sample_index = fr.p.subframe_index * fr.seed_stride + fr.seed_offset;   // = subframe_index in the reference
    uint32_t seed = tea<4>((uint32_t)i, sample_index);
    float jx = 0.5f, jy = 0.5f;
    if (sample_index != 0) {
        jx = rnd(seed);
        jy = rnd(seed);
    }
    float3 ray_direction = camera_dir_exact(ld3(fr.p.U), ld3(fr.p.V), ld3(fr.p.W), x, y, W, H, jx, jy);
    float3 ray_origin = ld3(fr.p.eye);
    // PayloadRadiance (whitted.h:86-108)
    float3 result = f3(0.0f), throughput = f3(1.0f), currentResult = f3(0.0f), vis_A = f3(0.0f), vis_B = f3(0.0f);
    float prd_pdf = 0.f;
    int depth = 0;
    bool done = false;
    while (true) {
        TravRay r{ray_origin.x, ray_origin.y, ray_origin.z, ray_direction.x, ray_direction.y, ray_direction.z, SPC_SCENE_EPS, 1e16f};
        TravHit h;
        if (!traverse_bvh8<false, false>(fr.sc.nodes, fr.sc.tris, r, true, stack, kPtPixBlock, h, cn, ct, s_lut)) {
            done = true;   // __miss__constant_radiance (environment lighting is out of scope)
            currentResult = f3(0.0f);
        } else {
            const LocalGeom geom = hit_geometry(fr.sc, h.prim, h.u, h.v);
            if (geom.light >= 0) {   // __closesthit__lightsource
                LightSample ls;
                light_reverse_sample(fr, geom.light, geom.uv.x, geom.uv.y, ls);
                if (dot(ray_direction, ls.normal) <= 0) {
                    float MIS_weight = 1;
                    if (depth != 0) {
                        const float pdf_hit = prd_pdf * fabsf(dot(ray_direction, ls.normal)) / (h.t * h.t);
                        const float pdf_area = ls.pdf;
                        MIS_weight = pdf_hit / (pdf_area + pdf_hit);
                    }
                    result += throughput * ls.emission * MIS_weight;
                }
                done = true;
            } else {   // __closesthit__radiance
                const Pbr pbr = shade_pbr(fr.sc, geom.material, geom.uv);
                float3 N = geom.Ng;
                if (dot(N, ray_direction) > 0.f) N = -N;
                const float3 in_dir = -ray_direction;
                float3 res = f3(0.0f);
                const float rr_rate = fmaxf(0.3f, fminf(fmax3(pbr.base_color), 1.0f));   // clamp(fmaxf(color), MIN_RR_RATE, 1.0)
                const int light_id = pick_light(fr, seed);
                LightSample ls;
                {
                    const float r1 = rnd(seed);
                    const float r2 = rnd(seed);
                    light_reverse_sample(fr, light_id, r1, r2, ls);
                }
                const float L_dist = length(ls.position - geom.P);
                const float3 L = (ls.position - geom.P) / L_dist;
                const float3 V = -normalize(ray_direction);
                const float L_dot_LN = dot(-L, ls.normal);
                const float N_dot_L = dot(N, L);
                const float N_dot_V = dot(N, V);
                if (N_dot_L > 0.0f && N_dot_V > 0.0f && L_dot_LN > 0.0f) {
                    vis_A = geom.P;
                    vis_B = ls.position;
                    const float3 eval = bsdf_eval(pbr, N, V, L);
                    const float pdf_area = ls.pdf;
                    const float pdf_hit = bsdf_pdf(pbr, N, V, L) * fabsf(L_dot_LN) / (L_dist * L_dist) * rr_rate;
                    const float MIS_weight = pdf_area / (pdf_hit + pdf_area);
                    res += throughput * ls.emission * 1.0f / ls.pdf * N_dot_L * L_dot_LN / L_dist / L_dist * eval * MIS_weight;
                }
                currentResult += res;
                ray_origin = geom.P;
                if (rnd(seed) > rr_rate) {
                    done = true;
                } else {
                    ray_direction = bsdf_sample(pbr, N, in_dir, seed);
                    const float pdf = bsdf_pdf(pbr, N, in_dir, ray_direction);
                    if (pdf > 0.0f) {
                        throughput *= bsdf_eval(pbr, N, in_dir, ray_direction) * fabsf(dot(ray_direction, N)) / pdf / rr_rate;
                        prd_pdf = pdf * rr_rate;
                    } else {
                        done = true;
                    }
                }
            }
        }
        if (sum3(currentResult) > 0.0f) {
            // visibilityTest (cuProg.h:489-502)
            const float3 bias_pos = vis_B - vis_A;
            const float len = length(bias_pos);
            const float3 dir = bias_pos / len;
            TravRay sr{vis_A.x, vis_A.y, vis_A.z, dir.x, dir.y, dir.z, SPC_SCENE_EPS, len - SPC_SCENE_EPS};
            TravHit sh;
            if (!traverse_bvh8<true, false>(fr.sc.nodes, fr.sc.tris, sr, false, stack, kPtPixBlock, sh, cn, ct, s_lut)) result += currentResult;
            currentResult = f3(0.0f);
        }
        if (done || depth > 30) break;
        depth += 1;
    }
    float3 c = result;
    if (fr.p.subframe_index > 0) {
        const float t = 1.0f / (float)(int)(fr.p.subframe_index + 1);
        const spc_float4 prev = fr.p.accum_buffer[i];
        c = lerp3(f3(prev.x, prev.y, prev.z), c, t);
    }
    fr.p.accum_buffer[i] = spc_float4{c.x, c.y, c.z, 1.0f};
    if (fr.p.frame_buffer) {
        const float lum = 0.3f * c.x + 0.6f * c.y + 0.1f * c.z;
        const float s = 1.0f + 1 * lum / 1.5f;
        const float inv = 1.0f / s;
        const float3 v = f3(clampf(c.x * 1.0f * inv, 0.f, 1.f), clampf(c.y * 1.0f * inv, 0.f, 1.f), clampf(c.z * 1.0f * inv, 0.f, 1.f));
        fr.p.frame_buffer[i] = pt_quantize8(pt_to_srgb(v.x)) | (pt_quantize8(pt_to_srgb(v.y)) << 8) | (pt_quantize8(pt_to_srgb(v.z)) << 16) | (255u << 24);
    }
}

void launch_pt(Context& c, int width, int height) {
    NvtxRange range("spc: pt");
    SPC_REQUIRE(c.has_params, SPC_ERR_INVALID, "spc_launch: spc_set_params has not been called");
    SPC_REQUIRE(width > 0 && height > 0 && (unsigned)width == c.params.width && (unsigned)height == c.params.height, SPC_ERR_INVALID,
                "spc_launch(pt): launch size %dx%d differs from MyParams %ux%u", width, height, c.params.width, c.params.height);
    SPC_REQUIRE(c.params.accum_buffer, SPC_ERR_INVALID, "spc_launch(pt): MyParams::accum_buffer is null");
    SPC_REQUIRE(c.geom.n_lights > 0, SPC_ERR_NO_SCENE, "spc_launch(pt): the scene has no lights");
    const int n = width * height;
    const DevFrame fr = make_dev_frame(c);
    k_pt<<<(n + kPtPixBlock - 1) / kPtPixBlock, kPtPixBlock, 0, c.stream>>>(fr, n);
    SPC_CUDA(cudaGetLastError());
    c.launches++;
}

}  // namespace spc
