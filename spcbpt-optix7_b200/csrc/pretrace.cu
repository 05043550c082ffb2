// pretrace.cu -- NEE training-path tracer.  Replaces optixLaunch of __raygen__TrainData (raygen.cu:751-868) with
// PreTrace_buildPathInfo (:708-740), rr_acc_accept (:741-749) and the TrainData::nVertex(_device) algebra
// (optixPathTracer.h:266-324, cuProg.h:1124-1292).
//
// One lane per training path: a unidirectional eye path from a random pixel; at every surface vertex NEE to a
// uniformly sampled emitter point (shadow-ray tested), plus emitter hits at path length > 2; reservoir sampling
// keeps ONE complete path per lane, written as a pathInfo_sample and one pathInfo_node per split point
// (eye prefix A | light suffix B) with peak_pdf = pdf(A) * contribution(B).
// This pass runs once during preprocessing; lanes are independent, so one launch may carry any number of paths
// (the reference issues 10 000 per launch; its seeds are tea<4>(launch index, iteration), reproduced here).
#include "shade.cuh"
#include "traverse.cuh"

namespace spc {

struct NVertex {   // TrainData::nVertex, optixPathTracer.h:266-324
    float3 position, dir, normal, weight, color;
    float  pdf;
    int    materialId, label_id, depth;
    bool   isBrdf;
    __device__ bool isLightSource() const { return materialId < 0; }
    __device__ bool isAreaLight() const { return materialId == -1; }
};

__device__ __forceinline__ NVertex nvertex_from(const Vtx& a, bool eye_side) {
    NVertex n;
    n.position = a.position; n.normal = a.normal; n.color = a.color;
    n.materialId = a.materialId; n.pdf = a.pdf; n.label_id = a.subspaceId; n.isBrdf = a.isBrdf != 0; n.depth = a.depth;
    n.dir = a.depth == 0 ? f3(0.0f) : normalize(a.lastPosition - a.position);
    n.weight = eye_side ? f3(n.pdf) : a.flux;
    if (!eye_side && a.depth == 0 && a.type == SPC_VTYPE_QUAD) n.materialId = -1;   // setLightSourceFlag(false)
    return n;
}
__device__ __forceinline__ Pbr nv_mat(const DevFrame& fr, const NVertex& v) {
    Pbr m = load_pbr(fr.sc, v.materialId);
    m.base_color = v.color;
    return m;
}
__device__ __forceinline__ float nv_forward_light_pdf(const DevFrame& fr, const NVertex& self, const NVertex& b) {   // cuProg.h:1190-1214
    const float3 vec = b.position - self.position;
    const float3 c_dir = normalize(vec);
    float g = fabsf(dot(c_dir, b.normal)) / dot(vec, vec);
    if (self.isLightSource()) {
        g *= fabsf(dot(self.normal, c_dir));
        return (float)((double)(self.pdf * g) * 1.0 / SPC_PI_D);
    }
    const Pbr mat = nv_mat(fr, self);
    const float d_pdf = bsdf_pdf(mat, self.normal, self.dir, c_dir);
    const float RR_rate = fmaxf(fmax3(self.color), 0.3f);
    return self.pdf * d_pdf * RR_rate * g;
}
__device__ __forceinline__ float3 nv_forward_eye(const DevFrame& fr, const NVertex& self, const NVertex& b) {   // cuProg.h:1216-1239
    const float3 vec = b.position - self.position;
    const float3 c_dir = normalize(vec);
    const float g = fabsf(dot(c_dir, b.normal)) / dot(vec, vec);
    const Pbr mat = nv_mat(fr, self);
    const float d_pdf = bsdf_pdf(mat, self.normal, self.dir, c_dir);
    const float RR_rate = fmaxf(fmax3(self.color), 0.3f);
    return self.weight * d_pdf * RR_rate * g;
}
__device__ __forceinline__ float3 nv_forward_light(const DevFrame& fr, const NVertex& self, const NVertex& b) {   // cuProg.h:1241-1281
    const float3 vec = b.position - self.position;
    const float3 c_dir = normalize(vec);
    if (self.isAreaLight()) {
        const float g = fabsf(dot(c_dir, b.normal)) * fabsf(dot(c_dir, self.normal)) / dot(vec, vec);
        return self.weight * g;
    }
    const float g = self.isBrdf ? fabsf(dot(c_dir, b.normal)) / dot(vec, vec)
                                : fabsf(dot(c_dir, b.normal)) * fabsf(dot(c_dir, self.normal)) / dot(vec, vec);
    const Pbr mat = nv_mat(fr, self);
    const float3 d_contri = bsdf_eval(mat, self.normal, self.dir, c_dir);
    return self.weight * g * d_contri;
}
__device__ __forceinline__ float3 nv_local_contri(const DevFrame& fr, const NVertex& self, const NVertex& b) {   // cuProg.h:1282-1290
    const float3 c_dir = normalize(b.position - self.position);
    const Pbr mat = nv_mat(fr, self);
    return bsdf_eval(mat, self.normal, self.dir, c_dir);
}
__device__ __forceinline__ NVertex nv_extend(const DevFrame& fr, const NVertex& a, const NVertex& b, bool eye_side) {   // nVertex_device(a,b,eye_side)
    NVertex n;
    n.position = a.position;
    n.dir = normalize(b.position - a.position);
    n.normal = a.normal;
    n.weight = eye_side ? nv_forward_eye(fr, b, a) : nv_forward_light(fr, b, a);
    n.pdf = eye_side ? n.weight.x : nv_forward_light_pdf(fr, b, a);
    n.color = a.color;
    n.materialId = a.materialId;
    n.label_id = a.label_id;
    n.isBrdf = a.isBrdf;
    n.depth = b.depth + 1;
    return n;
}
__device__ __forceinline__ void write_conn(spc_train_conn* c, const NVertex& a, const NVertex& b) {   // pathInfo_node(a, b)
    st3(c->A_position, a.position); st3(c->B_position, b.position);
    st3(c->A_dir, a.dir); st3(c->B_dir, b.dir);
    st3(c->A_normal, a.normal); st3(c->B_normal, b.normal);
    c->peak_pdf = a.weight.x * sum3(b.weight) * (float)(b.isBrdf ? 0 : 1) * (float)(a.isBrdf ? 0 : 1);
    c->path_id = 0;
    c->label_A = a.depth;   // set_eye_depth
    c->label_B = b.label_id;
    c->valid = 1;
    c->light_source = b.isLightSource() ? 1 : 0;
    c->_pad[0] = c->_pad[1] = 0;
}

constexpr int kMaxTrainVerts = 16;   // >= PRETRACE_CONN_PADDING (optixPathTracer.h:75)

// PreTrace_buildPathInfo (raygen.cu:708-740); `buf` = the eye vertices 0..n_buf-1, the path ends at buf[n_buf-1]
__device__ __noinline__ void build_path_info(const DevFrame& fr, const spc_vertex* buf, int n_buf, NVertex light, spc_train_path* path, spc_train_conn* conn) {
    int e = n_buf - 1;
    const int end_ind = n_buf - 1;
    Vtx eye = vtx_load(buf + e);
    NVertex n_eye = nvertex_from(eye, true);
    const NVertex n_next_eye = nv_extend(fr, light, n_eye, true);
    const float3 seg_contri = nv_local_contri(fr, n_eye, light);
    float sample_pdf = n_next_eye.pdf;
    sample_pdf += n_eye.pdf * light.pdf;
    float3 contri = eye.flux * nv_forward_light(fr, light, n_eye) * seg_contri;
    for (int i = 0; i < end_ind; i++) {
        write_conn(conn + (end_ind - i - 1), n_eye, light);
        e--;
        light = nv_extend(fr, n_eye, light, false);
        eye = vtx_load(buf + e);
        n_eye = nvertex_from(eye, true);
    }
    const float weight = sum3(contri) / sample_pdf;
    if (isnan(weight)) contri = f3(0.0f);
    if (isinf(weight)) contri = f3(0.0f);
    st3(path->contri, contri);
    path->sample_pdf = sample_pdf;
    path->fix_pdf = n_next_eye.pdf;
    path->begin_ind = 0;
    path->end_ind = end_ind;
    path->valid = 1;
}
__device__ __forceinline__ bool rr_acc_accept(int acc_num, uint32_t& seed) {   // raygen.cu:741-749
    const float r = rnd(seed);
    return 1.0f / (float)(acc_num + 1) > r;
}

constexpr int kPtBlock = 64;

__global__ void __launch_bounds__(kPtBlock) k_pretrace(const DevFrame fr, spc_vertex* __restrict__ scratch) {
    __shared__ uint2 s_stack[kSmStack * kPtBlock];
    __shared__ TravLut s_lut;
    trav_lut_init(s_lut);   // before any thread leaves
    const spc_pretrace_params& pt = fr.p.pre_tracer;
    const int launch_index = blockIdx.x * kPtBlock + threadIdx.x;
    if (launch_index >= pt.num_core) return;
    uint2* stack = s_stack + threadIdx.x;
    unsigned cn = 0, ct = 0;
    uint32_t seed = tea<4>((uint32_t)launch_index, (uint32_t)pt.iteration);
    const float jx = rnd(seed);   // make_float2(rnd(seed), rnd(seed)): nvcc evaluates left to right
    const float jy = rnd(seed);
    const float dx = __fsub_rn(__fmul_rn(2.0f, jx), 1.0f), dy = __fsub_rn(__fmul_rn(2.0f, jy), 1.0f);
    float3 ray_direction = pixel_dir_exact(dx, dy, ld3(fr.p.U), ld3(fr.p.V), ld3(fr.p.W));
    float3 ray_origin = ld3(fr.p.eye);
    spc_vertex* buffer = scratch + (size_t)launch_index * kMaxTrainVerts;   // BDPTVertex buffer[PRETRACE_CONN_PADDING]
    int buffer_size = 0;
    int resample_number = 0;
    Vtx cur;
    vtx_zero(cur);
    cur.position = ray_origin;
    cur.flux = f3(1.0f);
    cur.pdf = 1.0f;
    cur.RMIS_pointer = 0;
    cur.normal = ray_direction;
    cur.isOrigin = 1;
    cur.depth = 0;
    cur.singlePdf = 1.0f;
    float3 pre_flux = f3(0.f);
    float pre_singlePdf = 1.0f;
    const unsigned bufferBias = (unsigned)launch_index * (unsigned)pt.padding;
    spc_train_path* path = (spc_train_path*)pt.paths + launch_index;
    spc_train_conn* conn = (spc_train_conn*)pt.conns + bufferBias;
    spc_train_path pinit;
    pinit.contri = spc_float3{0.f, 0.f, 0.f};
    pinit.sample_pdf = 0.f; pinit.fix_pdf = 0.f; pinit.begin_ind = 0; pinit.end_ind = 0; pinit.choice_id = 0; pinit.pixel_x = 0; pinit.pixel_y = 0;
    pinit.valid = 0;
    for (int k = 0; k < 7; k++) pinit._pad[k] = 0;
    *path = pinit;
    vtx_store(buffer + buffer_size, cur);
    buffer_size++;
    bool done = false;
    int depth = 0;
    while (true) {
        TravRay r{ray_origin.x, ray_origin.y, ray_origin.z, ray_direction.x, ray_direction.y, ray_direction.z, SPC_SCENE_EPS, 1e16f};
        TravHit h;
        if (!traverse_bvh8<false, false>(fr.sc.nodes, fr.sc.tris, r, true, stack, kPtBlock, h, cn, ct, s_lut)) break;   // miss: no vertex
        const LocalGeom g = hit_geometry(fr.sc, h.prim, h.u, h.v);
        Vtx mid;
        if (g.light >= 0) {
            if (!eye_hits_light(fr, cur, pre_flux, pre_singlePdf, g, h.t, ray_direction, mid)) break;   // emitter seen from behind
            // path.size counts the camera vertex: size > 2 <=> at least one surface vertex before the emitter
            if (buffer_size + 1 > 2 && rr_acc_accept(resample_number, seed)) {
                LightSample ls;
                light_reverse_sample(fr, mid.materialId, mid.uv.x, mid.uv.y, ls);
                Vtx lv;
                vtx_zero(lv);
                init_vertex_from_light_sample(ls, lv);
                build_path_info(fr, buffer, buffer_size, nvertex_from(lv, false), path, conn);
                resample_number++;
            }
            break;
        }
        SurfaceOut so;
        surface_hit(fr, cur, pre_flux, pre_singlePdf, g, h.t, ray_direction, false, seed, mid, so);
        cur = mid;
        pre_flux = so.next_flux;
        pre_singlePdf = so.next_singlePdf;
        done = so.done;
        vtx_store(buffer + buffer_size, cur);
        buffer_size++;
        // NEE: lightSample(seed) = pick + position (cuProg.h:622-626), shadow ray, reservoir acceptance
        LightSample ls;
        {
            const int li = pick_light(fr, seed);
            const float r1 = rnd(seed);
            const float r2 = rnd(seed);
            light_reverse_sample(fr, li, r1, r2, ls);
        }
        const float3 vis_vec = ls.position - cur.position;
        bool visible;
        {
            const float len = length(vis_vec);
            const float3 dir = vis_vec / len;
            TravRay sr{cur.position.x, cur.position.y, cur.position.z, dir.x, dir.y, dir.z, SPC_SCENE_EPS, len - SPC_SCENE_EPS};
            TravHit sh;
            visible = !traverse_bvh8<true, false>(fr.sc.nodes, fr.sc.tris, sr, false, stack, kPtBlock, sh, cn, ct, s_lut);
        }
        if (visible && rr_acc_accept(resample_number, seed)) {
            if (dot(vis_vec, ls.normal) < 0) {
                Vtx lv;
                vtx_zero(lv);
                init_vertex_from_light_sample(ls, lv);
                build_path_info(fr, buffer, buffer_size, nvertex_from(lv, false), path, conn);
                resample_number++;
            }
        }
        if (done || depth > fr.max_depth) break;
        if (buffer_size >= pt.padding) break;
        ray_direction = so.dir;
        ray_origin = g.P;
        depth += 1;
    }
    int beginIndex = 0;
    const bool valid = path->valid != 0;
    if (valid) beginIndex += path->end_ind - path->begin_ind;
    for (int i = beginIndex; i < pt.padding; i++) {
        conn[i].valid = 0;
        conn[i].light_source = 0;
    }
    path->sample_pdf = path->sample_pdf / (float)resample_number;
    path->begin_ind += (int)bufferBias;
    path->end_ind += (int)bufferBias;
    path->pixel_x = (int)((float)fr.p.width * jx);
    path->pixel_y = (int)((float)fr.p.height * jy);
    if (path->begin_ind == path->end_ind && valid) path->valid = 0;
}

DevFrame make_dev_frame(Context& c);

void launch_pretrace(Context& c) {
    NvtxRange range("spc: pretrace");
    SPC_REQUIRE(c.has_params, SPC_ERR_INVALID, "spc_launch: spc_set_params has not been called");
    const spc_pretrace_params& pt = c.params.pre_tracer;
    SPC_REQUIRE(pt.num_core > 0 && pt.padding > 0 && pt.padding <= kMaxTrainVerts && pt.paths && pt.conns, SPC_ERR_INVALID,
                "spc_launch(pretrace): MyParams::pre_tracer is not set up (padding <= %d)", kMaxTrainVerts);
    SPC_REQUIRE(c.geom.n_lights > 0, SPC_ERR_NO_SCENE, "spc_launch(pretrace): the scene has no lights");
    c.pretrace_scratch.alloc((size_t)pt.num_core * kMaxTrainVerts);
    const DevFrame fr = make_dev_frame(c);
    k_pretrace<<<(pt.num_core + kPtBlock - 1) / kPtBlock, kPtBlock, 0, c.stream>>>(fr, c.pretrace_scratch.p);
    SPC_CUDA(cudaGetLastError());
    c.launches++;
}

}  // namespace spc
