// orc_train.cpp -- CPU oracle of the subspace-training path.  TEST INFRASTRUCTURE ONLY.
// Restates (paths under /root/reference/src/OptiXPathTracer):
//   * __raygen__TrainData, PreTrace_buildPathInfo, rr_acc_accept            raygen.cu:708-868
//   * TrainData::nVertex / nVertex_device / pathInfo_node / pathInfo_sample  optixPathTracer.h:266-383, cuProg.h:1124-1292
//   * MyThrustOp::valid_sample_gather, sample_reweight, get_weighted_point_for_tree_building, node_label,
//     preprocess_getQ, Q_zero_handle, build_optimal_E_train_data, preprocess_getGamma, train_optimal_E,
//     Gamma2CMFGamma                                                         cuda_thrust/device_thrust.cu
// The raygen part is pinned bit-for-bit against the reference's own program on the host shim
// (tests/test_oracle_vs_ref.py); the MyThrustOp part against the reference's own library on the GPU box
// (oracle/_ref/libref_thrust.so, tests/test_train_gpu.py).
#include <algorithm>
#include <cfloat>
#include <cstring>
#include "orc_internal.h"
#include "orc_train.h"

namespace orc {

static const float PIf = 3.14159265358979323846f;
static const double PId = 3.14159265358979323846;
static inline float absf(float x) { return std::fabs(x); }

// TrainData::nVertex (optixPathTracer.h:266-324) + nVertex_device (cuProg.h:1128-1292)
struct NVertex {
    f3 position, dir, normal, weight, color;
    float pdf;
    int materialId, label_id, depth;
    bool isBrdf;
    bool isLightSource() const { return materialId < 0; }
    bool isAreaLight() const { return materialId == -1; }
};

static NVertex nvertex_from(const spc_vertex& a, bool eye_side) {   // nVertex(const BDPTVertex&, bool)
    NVertex n;
    n.position = ld(a.position); n.normal = ld(a.normal); n.color = ld(a.color);
    n.materialId = a.materialId; n.pdf = a.pdf; n.label_id = a.subspaceId; n.isBrdf = a.isBrdf != 0; n.depth = a.depth;
    n.dir = a.depth == 0 ? mk3(0.0f) : normalize(ld(a.lastPosition) - ld(a.position));
    n.weight = eye_side ? mk3(n.pdf) : ld(a.flux);
    if (!eye_side && a.depth == 0 && a.type == SPC_VTYPE_QUAD) n.materialId = -1;   // setLightSourceFlag(false)
    return n;
}
static Pbr nv_mat(const Frame& fr, const NVertex& v) {
    Pbr m = load_pbr(*fr.sc, v.materialId);
    m.base_color = v.color;
    return m;
}
static float nv_forward_light_pdf(const Frame& fr, const NVertex& self, const NVertex& b) {   // cuProg.h:1190-1214
    const f3 vec = b.position - self.position;
    const f3 c_dir = normalize(vec);
    float g = absf(dot(c_dir, b.normal)) / dot(vec, vec);
    if (self.isLightSource()) {
        g *= absf(dot(self.normal, c_dir));
        return (float)((double)(self.pdf * g) * 1.0 / PId);
    }
    const Pbr mat = nv_mat(fr, self);
    const float d_pdf = bsdf_pdf(mat, self.normal, self.dir, c_dir);
    const float RR_rate = std::fmax(fmax3(self.color), 0.3f);
    return self.pdf * d_pdf * RR_rate * g;
}
static f3 nv_forward_eye(const Frame& fr, const NVertex& self, const NVertex& b) {   // cuProg.h:1216-1239
    const f3 vec = b.position - self.position;
    const f3 c_dir = normalize(vec);
    const float g = absf(dot(c_dir, b.normal)) / dot(vec, vec);
    const Pbr mat = nv_mat(fr, self);
    const float d_pdf = bsdf_pdf(mat, self.normal, self.dir, c_dir);
    const float RR_rate = std::fmax(fmax3(self.color), 0.3f);
    return self.weight * d_pdf * RR_rate * g;
}
static f3 nv_forward_light(const Frame& fr, const NVertex& self, const NVertex& b) {   // cuProg.h:1241-1281
    const f3 vec = b.position - self.position;
    const f3 c_dir = normalize(vec);
    if (self.isAreaLight()) {
        const float g = absf(dot(c_dir, b.normal)) * absf(dot(c_dir, self.normal)) / dot(vec, vec);
        return self.weight * g;
    }
    const float g = self.isBrdf ? absf(dot(c_dir, b.normal)) / dot(vec, vec)
                                : absf(dot(c_dir, b.normal)) * absf(dot(c_dir, self.normal)) / dot(vec, vec);
    const Pbr mat = nv_mat(fr, self);
    const f3 d_contri = bsdf_eval(mat, self.normal, self.dir, c_dir);
    return self.weight * g * d_contri;
}
static f3 nv_local_contri(const Frame& fr, const NVertex& self, const NVertex& b) {   // cuProg.h:1282-1290
    const f3 c_dir = normalize(b.position - self.position);
    const Pbr mat = nv_mat(fr, self);
    return bsdf_eval(mat, self.normal, self.dir, c_dir);
}
static NVertex nv_extend(const Frame& fr, const NVertex& a, const NVertex& b, bool eye_side) {   // nVertex_device(a, b, eye_side), cuProg.h:1130-1148
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
static void make_conn(spc_train_conn& c, const NVertex& a, const NVertex& b) {   // pathInfo_node(a, b), optixPathTracer.h:343-352
    memset(&c, 0, sizeof(c));
    st(c.A_position, a.position); st(c.B_position, b.position);
    st(c.A_dir, a.dir); st(c.B_dir, b.dir);
    st(c.A_normal, a.normal); st(c.B_normal, b.normal);
    c.valid = 1;
    c.light_source = b.isLightSource() ? 1 : 0;
    c.label_B = b.label_id;
    c.peak_pdf = a.weight.x * sum3(b.weight) * (float)(b.isBrdf ? 0 : 1) * (float)(a.isBrdf ? 0 : 1);
    c.label_A = a.depth;   // set_eye_depth
}

// PreTrace_buildPathInfo (raygen.cu:708-740)
static void build_path_info(const Frame& fr, const spc_vertex* eye, NVertex light, spc_train_path* path, spc_train_conn* conn, int pathSize) {
    path->valid = 1;
    path->begin_ind = 0;
    path->end_ind = pathSize - 1;
    path->sample_pdf = 0;
    NVertex n_eye = nvertex_from(*eye, true);
    const NVertex n_next_eye = nv_extend(fr, light, n_eye, true);
    const f3 seg_contri = nv_local_contri(fr, n_eye, light);
    path->sample_pdf = n_next_eye.pdf;
    path->sample_pdf += n_eye.pdf * light.pdf;
    path->fix_pdf = n_next_eye.pdf;
    st(path->contri, ld(eye->flux) * nv_forward_light(fr, light, n_eye) * seg_contri);
    for (int i = 0; i < path->end_ind; i++) {
        make_conn(conn[path->end_ind - i - 1], n_eye, light);
        eye--;
        light = nv_extend(fr, n_eye, light, false);
        n_eye = nvertex_from(*eye, true);
    }
    const float weight = sum3(ld(path->contri)) / path->sample_pdf;
    if (std::isnan(weight)) st(path->contri, mk3(0.0f));
    if (std::isinf(weight)) st(path->contri, mk3(0.0f));
}
static bool rr_acc_accept(int acc_num, uint32_t& seed) {   // raygen.cu:741-749
    const float r = rnd(seed);
    return 1.0f / (float)(acc_num + 1) > r;
}

// __raygen__TrainData (raygen.cu:751-868)
void pretrace_core(const Frame& fr, int launch_index) {
    const spc_params& P = fr.p;
    const spc_pretrace_params& pt = P.pre_tracer;
    uint32_t seed = tea(4, (uint32_t)launch_index, (uint32_t)pt.iteration);
    float jx, jy;
    if (g_jitter_rtl) { jy = rnd(seed); jx = rnd(seed); }
    else { jx = rnd(seed); jy = rnd(seed); }
    const float dx = 2.0f * jx - 1.0f, dy = 2.0f * jy - 1.0f;
    f3 ray_direction = normalize(dx * ld(P.U) + dy * ld(P.V) + ld(P.W));
    f3 ray_origin = ld(P.eye);
    spc_vertex buffer[16];
    int buffer_size = 0;
    int resample_number = 0;
    Payload payload;
    memset(&payload.path, 0, sizeof(payload.path));
    for (int k = 0; k < 3; k++) payload.path.v[k].type = SPC_VTYPE_QUAD;
    payload.clear();
    payload.seed = seed;
    payload.ray_direction = ray_direction;
    payload.origin = ray_origin;
    {   // init_EyeSubpath (raygen.cu:216-231)
        payload.path.size++;
        spc_vertex& v = payload.path.cur();
        st(v.position, ray_origin);
        st(v.flux, mk3(1.0f));
        v.pdf = 1.0f;
        v.RMIS_pointer = 0;
        st(v.normal, ray_direction);
        v.isOrigin = 1;
        v.depth = 0;
        v.singlePdf = 1.0f;
        payload.path.next().singlePdf = 1.0f;
    }
    const unsigned bufferBias = (unsigned)launch_index * (unsigned)pt.padding;
    spc_train_path* currentPath = (spc_train_path*)pt.paths + launch_index;
    spc_train_conn* currentConn = (spc_train_conn*)pt.conns + bufferBias;
    memset(currentPath, 0, sizeof(*currentPath));
    currentPath->valid = 0;
    buffer[buffer_size++] = payload.path.cur();
    while (true) {
        const int begin_depth = payload.path.size;
        trace_subpath(fr, payload, ray_origin, ray_direction, false);
        if (payload.path.size == begin_depth) break;
        if (payload.path.cur().type == SPC_VTYPE_HIT_LIGHT_SOURCE) {
            if (payload.path.size > 2 && rr_acc_accept(resample_number, payload.seed)) {
                const spc_vertex& cur = payload.path.cur();
                LightSample ls;
                light_reverse_sample(fr, fr.sc->lights[cur.materialId], cur.uv.x, cur.uv.y, ls);
                spc_vertex lv;
                memset(&lv, 0, sizeof(lv));
                lv.type = SPC_VTYPE_QUAD;
                init_vertex_from_light_sample(ls, lv);
                build_path_info(fr, buffer + buffer_size - 1, nvertex_from(lv, false), currentPath, currentConn, buffer_size);
                resample_number++;
            }
            break;
        }
        buffer[buffer_size++] = payload.path.cur();
        const spc_vertex& eye_subpath = payload.path.cur();
        LightSample ls;
        light_sample_pos(fr, fr.sc->lights[pick_light(fr, payload.seed)], payload.seed, ls);
        const f3 vis_vec = ls.position - ld(eye_subpath.position);
        spc_vertex lv;
        memset(&lv, 0, sizeof(lv));
        lv.type = SPC_VTYPE_QUAD;
        init_vertex_from_light_sample(ls, lv);
        if (visibility_test(fr, ld(eye_subpath.position), ld(lv.position)) && rr_acc_accept(resample_number, payload.seed)) {
            if (dot(vis_vec, ld(ls.light->normal)) < 0) {
                build_path_info(fr, buffer + buffer_size - 1, nvertex_from(lv, false), currentPath, currentConn, buffer_size);
                resample_number++;
            }
        }
        if (payload.done || payload.depth > fr.max_depth) break;
        if (buffer_size >= pt.padding) break;
        ray_direction = payload.ray_direction;
        ray_origin = payload.origin;
        payload.depth += 1;
    }
    int beginIndex = 0;
    if (currentPath->valid) beginIndex += currentPath->end_ind - currentPath->begin_ind;
    for (int i = beginIndex; i < pt.padding; i++) {
        memset(&currentConn[i], 0, sizeof(spc_train_conn));
        currentConn[i].valid = 0;
    }
    currentPath->sample_pdf /= (float)resample_number;
    currentPath->begin_ind += (int)bufferBias;
    currentPath->end_ind += (int)bufferBias;
    currentPath->pixel_x = (int)((float)P.width * jx);
    currentPath->pixel_y = (int)((float)P.height * jy);
    if (currentPath->begin_ind == currentPath->end_ind && currentPath->valid) currentPath->valid = 0;
}

// =============================================================================================
// MyThrustOp seam: the reference does most of this in serial host loops after full D2H copies
// =============================================================================================
// valid_sample_gather (device_thrust.cu:457-493): order-preserving compaction + index fix-up, appended to the set
int TrainSet::gather(const spc_train_path* raw_paths, int max_paths, const spc_train_conn* raw_conns, int max_conns) {
    std::vector<int> flag(max_conns);   // exclusive scan of the conn valid flags
    int run = 0;
    for (int i = 0; i < max_conns; i++) { flag[i] = run; run += raw_conns[i].valid ? 1 : 0; }
    const int node_bias = (int)conns.size(), sample_bias = (int)paths.size();
    int sample_count = 0;
    for (int i = 0; i < max_conns; i++) if (raw_conns[i].valid) conns.push_back(raw_conns[i]);
    for (int i = 0; i < max_paths; i++) {
        if (!raw_paths[i].valid) continue;
        spc_train_path s = raw_paths[i];
        const int bias = s.begin_ind - flag[s.begin_ind];
        s.begin_ind += node_bias - bias;
        s.end_ind += node_bias - bias;
        for (int k = s.begin_ind; k < s.end_ind; k++) conns[k].path_id = sample_count + sample_bias;
        paths.push_back(s);
        sample_count++;
    }
    return sample_count;
}

// sample_reweight (device_thrust.cu:574-623): per 10x10-pixel block, divide the contribution by mean-ish weight
void TrainSet::reweight() {
    const int nb = (int)(1920 * 1000 / 100 * 1.1);
    std::vector<float> weight(nb, 0.f);
    for (size_t i = 0; i < paths.size(); i++) {
        const int n_id = paths[i].pixel_x / 10 + paths[i].pixel_y / 10 * 192;
        const float ww = sum3(ld(paths[i].contri)) / paths[i].sample_pdf;
        if (std::isnan(ww) || std::isinf(ww)) continue;
        weight[n_id] += ww;
    }
    for (size_t i = 0; i < paths.size(); i++) {
        const int n_id = paths[i].pixel_x / 10 + paths[i].pixel_y / 10 * 192;
        const float w = (float)((double)(weight[n_id] / 100) + 0.1);
        st(paths[i].contri, ld(paths[i].contri) / w);
    }
}

// get_weighted_point_for_tree_building (device_thrust.cu:494-527).  Light side: the reference pushes an
// UNINITIALISED divide_weight for connections whose light endpoint is an emitter; here they are zero-weight
// samples at the origin (documented deviation: the reference's values are indeterminate).
std::vector<spc_divide_weight> TrainSet::tree_points(bool eye_side, int max_size) const {
    std::vector<spc_divide_weight> ans;
    const size_t limit = max_size == 0 ? paths.size() : std::min(paths.size(), (size_t)max_size);
    for (size_t i = 0; i < limit; i++)
        for (int j = paths[i].begin_ind; j < paths[i].end_ind; j++) {
            spc_divide_weight t;
            memset(&t, 0, sizeof(t));
            float w = sum3(ld(paths[i].contri)) / paths[i].sample_pdf;
            if (std::isnan(w) || std::isinf(w)) w = 0.f;   // documented deviation, see csrc/train.cu k_tree_points
            if (eye_side) {
                t.dir = conns[j].A_dir; t.normal = conns[j].A_normal; t.position = conns[j].A_position;
                t.weight = w;
            } else if (!conns[j].light_source) {
                t.dir = conns[j].B_dir; t.normal = conns[j].B_normal; t.position = conns[j].B_position;
                t.weight = w;
            }
            ans.push_back(t);
        }
    return ans;
}

// node_label (device_thrust.cu:554-573)
void TrainSet::label(const spc_tree_node* eye_tree, const spc_tree_node* light_tree) {
    for (auto& s : conns) {
        s.label_A = tree_label(eye_tree, ld(s.A_position), ld(s.A_normal));
        if (!s.light_source) s.label_B = tree_label(light_tree, ld(s.B_position), ld(s.B_normal));
    }
}

// preprocess_getQ (device_thrust.cu:347-409): running mean of per-launch Q estimates; returns the accumulated path count
int QEstimator::add(const spc_vertex* lvc, const uint8_t* valid, int n) {
    std::vector<float> tmp(K, 0.f);
    int path_count = 0;
    for (int i = 0; i < n; i++) if (valid[i] && lvc[i].depth == 0) path_count++;
    acc_valid_path += path_count;
    const float t = (float)path_count / (float)acc_valid_path;
    for (int i = 0; i < n; i++) {
        if (!valid[i]) continue;
        float res = (lvc[i].flux.x + lvc[i].flux.y + lvc[i].flux.z) / lvc[i].pdf;
        res = std::isinf(res) ? 0 : res;
        const float w = std::isnan(res) ? 0 : res;
        tmp[lvc[i].subspaceId] += w;
    }
    for (int i = 0; i < K; i++) {
        tmp[i] /= (float)path_count;
        Q[i] = Q[i] * (1 - t) + tmp[i] * t;
    }
    return acc_valid_path;
}
void QEstimator::zero_handle() {   // Q_zero_handle (device_thrust.cu:335-346)
    for (int i = 0; i < K; i++) if (Q[i] == 0) Q[i] = FLT_MAX;
}

// build_optimal_E_train_data (device_thrust.cu:3261-3325) with its functors (:3114-3258)
static const float kLossThreshold = 1000000.0f;   // optimal_E_loss_threshold
static float outlier_value(const TrainSet& ts, const spc_train_path& s, const float* Q) {   // get_outler_value
    float outler_value = s.fix_pdf;
    const float weight = sum3(ld(s.contri));
    float loss = weight * weight / s.sample_pdf;
    if (loss > kLossThreshold || std::isnan(loss)) loss = kLossThreshold;
    for (int i = s.begin_ind; i < s.end_ind; i++) outler_value = (float)((double)outler_value + (double)(ts.conns[i].peak_pdf / Q[ts.conns[i].label_B]) / 1000.0);   // float += double
    return loss / outler_value;
}
void TrainSet::build_train_data(int n_samples, const float* Q, int K, TrainData& td) {
    td.N = n_samples;
    td.M = paths[n_samples - 1].end_ind;
    std::vector<float> t(1000);
    for (int i = 0; i < 1000; i++) t[i] = outlier_value(*this, paths[i], Q);
    std::sort(t.begin(), t.end());
    td.outlier_threshold = t[999];
    for (auto& s : paths)
        if (outlier_value(*this, s, Q) > td.outlier_threshold) st(s.contri, ld(s.contri) * 0.0f);
    td.f_square.resize(td.N); td.pdf0.resize(td.N); td.P2N.resize(td.N);
    td.peak.resize(td.M); td.label_E.resize(td.M); td.label_P.resize(td.M);
    for (int id = 0; id < td.N; id++) {   // construct_optimal_E_data_sample
        const spc_train_path& s = paths[id];
        const float weight = sum3(ld(s.contri));
        float f = weight * weight / s.sample_pdf;
        if (f > kLossThreshold || std::isnan(f)) f = kLossThreshold;
        td.f_square[id] = f;
        td.pdf0[id] = s.fix_pdf;
        td.P2N[id] = s.begin_ind;
    }
    for (int id = 0; id < td.M; id++) {   // construct_optimal_E_data_node
        const spc_train_conn& s = conns[id];
        td.label_E[id] = s.label_A * K + s.label_B;
        td.label_P[id] = s.path_id;
        float p = (double)Q[s.label_B] > 0.0 ? s.peak_pdf / Q[s.label_B] : 0.0f;
        if (std::isnan(p) || std::isinf(p)) p = 0;
        td.peak[id] = p;
    }
}

// preprocess_getGamma (device_thrust.cu:627-667)
void TrainSet::gamma_histogram(int K, std::vector<float>& G) const {
    G.assign((size_t)K * K, 0.f);
    for (size_t i = 0; i < paths.size(); i++) {
        const float weight = sum3(ld(paths[i].contri)) / paths[i].sample_pdf;
        for (int j = paths[i].begin_ind; j < paths[i].end_ind; j++) {
            const float weight2 = (float)std::fmin((double)weight, 10.0);   // CUDA host min(float,double) = fmin: NaN -> 10
            G[(size_t)conns[j].label_A * K + conns[j].label_B] += weight2;
        }
    }
    for (int i = 0; i < K; i++) {
        float weightS = 0;
        for (int j = 0; j < K; j++) weightS += G[(size_t)i * K + j];
        for (int j = 0; j < K; j++) {
            G[(size_t)i * K + j] /= weightS;
            if (weightS <= 1e-10f) G[(size_t)i * K + j] = (float)(1.0 / K);
        }
    }
}

// train_optimal_E (device_thrust.cu:3327-3344) = matrix_parameter::fit (:1615-1655) over matrix_optimal_operator
// (:923-1190) + Adam (:1437-1470).  thrust's reductions have no defined summation order, so this restatement
// (sequential sums) is a tolerance-level oracle for this stage, not a bit-level one.
void train_gamma(const TrainData& td, int K, std::vector<float>& G, int batch_size, int epochs, float lr, double conservative_d, std::vector<float>* loss_log) {
    const size_t n = (size_t)K * K;
    std::vector<float> theta(n), m(n, 0.f), v(n, 0.f), E(n), Esum(K), dE(n), dEsum(K), g(n);
    for (size_t i = 0; i < n; i++) theta[i] = (float)(-std::log(1.0 / (double)G[i] - 1));   // inver_sigmoid
    const float beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f;
    const float c_keep = (float)(1 - (double)conservative_d), c_uniform = (float)(conservative_d / (double)(float)K);   // constant_iterator<double> -> float
    int t = 0;
    const int num_batches = td.N / batch_size;
    std::vector<float> pdfs;
    for (int ep = 0; ep < epochs; ep++)
        for (int b = 0; b < num_batches; b++) {
            const int bs = b * batch_size;
            const int bn = td.P2N[bs];
            // the reference gives the last batch every remaining node (device_thrust.cu:1636), well-defined only when N is a
            // multiple of the batch size (its own use); here a batch always owns exactly the nodes of its paths
            const int seg = (bs + batch_size < td.N ? td.P2N[bs + batch_size] : td.M) - bn;
            // get_E
            for (int i = 0; i < K; i++) {
                float s = 0;
                for (int j = 0; j < K; j++) { E[(size_t)i * K + j] = (float)(1.0 / (1.0 + (double)cm_expf(-theta[(size_t)i * K + j]))); s += E[(size_t)i * K + j]; }
                Esum[i] = s;
                for (int j = 0; j < K; j++) E[(size_t)i * K + j] = E[(size_t)i * K + j] / s * c_keep + c_uniform;
            }
            // forward pdfs + loss gradient
            pdfs.assign(batch_size, 0.f);
            for (int k = 0; k < seg; k++) pdfs[td.label_P[bn + k] % batch_size] += td.peak[bn + k] * E[td.label_E[bn + k]];
            double loss = 0;
            for (int i = 0; i < batch_size; i++) {
                pdfs[i] += td.pdf0[bs + i];
                loss += td.f_square[bs + i] / pdfs[i];
                pdfs[i] = -td.f_square[bs + i] / pdfs[i] / pdfs[i];   // inver_gradient
            }
            if (loss_log) loss_log->push_back((float)(loss / batch_size));
            std::fill(dE.begin(), dE.end(), 0.f);
            for (int k = 0; k < seg; k++) dE[td.label_E[bn + k]] += td.peak[bn + k] * pdfs[td.label_P[bn + k] % batch_size];
            // gradient_E2theta
            for (int i = 0; i < K; i++) {
                float s = 0;
                for (int j = 0; j < K; j++) {
                    const float res = E[(size_t)i * K + j], den = Esum[i];
                    const float value = res * den;
                    s += (-value / den / den) * dE[(size_t)i * K + j];   // inver_gradient_res * dE
                }
                dEsum[i] = s;
            }
            for (size_t i = 0; i < n; i++) {
                const int row = (int)(i / K);
                const float sig = (float)(1.0 / (1.0 + (double)cm_expf(-theta[i])));
                const float a = sig * (1 - sig) * dEsum[row];                         // sigmoid_gradient_theta * dE_sum
                const float sg = E[i] * Esum[row];                                     // theta_gradient
                const float b0 = sg * (1 - sg) / Esum[row] * dE[i];
                g[i] = a + b0;
            }
            // Adam (adam_step_func)
            t += 1;
            for (size_t i = 0; i < n; i++) {
                m[i] = beta1 * m[i] + (1 - beta1) * g[i];
                v[i] = beta2 * v[i] + (1 - beta2) * (g[i] * g[i]);
                const float m_hat = m[i] / (1 - cm_powf(beta1, (float)t));
                const float v_hat = v[i] / (1 - cm_powf(beta2, (float)t));
                const float step = m_hat / (std::sqrt(v_hat) + eps);
                if (!std::isnan(step)) theta[i] -= lr * step;
            }
        }
    // toE: sigmoid + row normalise (no conservative mixing)
    for (int i = 0; i < K; i++) {
        float s = 0;
        for (int j = 0; j < K; j++) { G[(size_t)i * K + j] = (float)(1.0 / (1.0 + (double)cm_expf(-theta[(size_t)i * K + j]))); s += G[(size_t)i * K + j]; }
        for (int j = 0; j < K; j++) G[(size_t)i * K + j] /= s;
    }
}

// Gamma2CMFGamma (device_thrust.cu:3406-3433)
void gamma_to_cmf(const std::vector<float>& G, int K, float conservative, std::vector<float>& cmf) {
    cmf = G;
    for (size_t i = 0; i < cmf.size(); i++) cmf[i] = (float)((double)(cmf[i] * (1 - conservative)) + (1.0 / K) * (double)conservative);
    for (int i = 0; i < K; i++) {
        for (int j = 1; j < K; j++) cmf[(size_t)i * K + j] += cmf[(size_t)i * K + j - 1];
        cmf[(size_t)(i + 1) * K - 1] = 1;
    }
}

}  // namespace orc
