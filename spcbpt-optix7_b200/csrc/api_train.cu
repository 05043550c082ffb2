// api_train.cu -- extern "C" entry points of the subspace-training path (MyThrustOp seam, part 2) and the
// plain memory helpers.
#include <cstring>
#include <cstring>
#include <vector>
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

namespace spc {
// Installs a classification tree as the context's eye / light tree: validation, reference-layout device copy (uploaded unless it is
// already there: upload = false after a device-side build) and the compact copy for the device-side walks.
void tree_install(Context& c, int eye_side, const spc_tree_node* nodes_host, int n, bool upload) {
    // A malformed tree (hand-edited tree_*.txt, a state trained with another K) would send the device walks out of bounds or into a
    // cycle, and labels >= K index Q / the CMFGamma rows out of bounds: reject it here.  Well-formed = what the builder emits
    // (classTree_host.h:103-284 appends the 8 children of a split behind their parent): root at 0, every child index in (parent, n)
    // and referenced once, node types 0..2, leaf labels in [0, K).
    {
        std::vector<uint8_t> seen((size_t)n, 0);
        for (int i = 0; i < n; i++) {
            const spc_tree_node& nd = nodes_host[i];
            if (nd.leaf) {
                SPC_REQUIRE(nd.label >= 0 && nd.label < c.K, SPC_ERR_INVALID, "tree: node %d: leaf label %d outside [0, %d)", i, nd.label, c.K);
                continue;
            }
            SPC_REQUIRE(nd.type >= 0 && nd.type <= 2, SPC_ERR_INVALID, "tree: node %d: type %d", i, nd.type);
            for (int k = 0; k < 8; k++) {
                const int ch = nd.child[k];
                SPC_REQUIRE(ch > i && ch < n, SPC_ERR_INVALID, "tree: node %d: child %d = %d outside (%d, %d)", i, k, ch, i, n);
                SPC_REQUIRE(!seen[ch], SPC_ERR_INVALID, "tree: node %d is the child of two nodes", ch);
                seen[ch] = 1;
            }
        }
    }
    DevBuf<spc_tree_node>& b = eye_side ? c.train.eye_tree : c.train.light_tree;
    if (upload) {
        ctree_register(b.p, nullptr, &c);
        b.alloc(n);
        SPC_CUDA(cudaMemcpyAsync(b.p, nodes_host, (size_t)n * sizeof(spc_tree_node), cudaMemcpyHostToDevice, c.stream));
    }
    // compact copy for the device-side walks (shade.cuh "compact trees"): 48 B per node = {mid, type} + 8 children, a leaf child
    // carries its label in the parent's entry (0x80000000 | label), a leaf root in the root's type word
    std::vector<float> ct((size_t)n * 12, 0.f);
    for (int i = 0; i < n; i++) {
        const spc_tree_node& nd = nodes_host[i];
        uint32_t w[12] = {};
        memcpy(&w[0], &nd.mid.x, 4); memcpy(&w[1], &nd.mid.y, 4); memcpy(&w[2], &nd.mid.z, 4);
        if (nd.leaf) {
            w[3] = 0x80000000u | (uint32_t)nd.label;
        } else {
            w[3] = (uint32_t)nd.type;
            for (int k = 0; k < 8; k++) {
                const int ch = nd.child[k];
                w[4 + k] = nodes_host[ch].leaf ? (0x80000000u | (uint32_t)nodes_host[ch].label) : (uint32_t)ch;
            }
        }
        memcpy(&ct[(size_t)i * 12], w, sizeof(w));
    }
    DevBuf<float4>& cb = eye_side ? c.train.eye_ctree : c.train.light_ctree;
    cb.alloc((size_t)n * 3);
    SPC_CUDA(cudaMemcpyAsync(cb.p, ct.data(), ct.size() * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    ctree_register(b.p, cb.p, &c);
}
}  // namespace spc

extern "C" {

int spc_valid_sample_gather(spc_context* ctx, const spc_train_path* raw_paths_dev, int max_paths, const spc_train_conn* raw_conns_dev,
                            int max_conns, int* sample_count) {
    SPC_API_BEGIN
    SPC_REQUIRE(raw_paths_dev && raw_conns_dev && max_paths > 0 && max_conns > 0, SPC_ERR_INVALID, "spc_valid_sample_gather: bad arguments");
    const int n = spc::train_gather(c, raw_paths_dev, max_paths, raw_conns_dev, max_conns);
    if (sample_count) *sample_count = n;
    SPC_API_END
}
int spc_sample_reweight(spc_context* ctx) {
    SPC_API_BEGIN
    spc::train_reweight(c);
    SPC_API_END
}
int spc_get_tree_points(spc_context* ctx, int eye_side, int max_size, spc_divide_weight* out_host, int cap, int* n) {
    SPC_API_BEGIN
    SPC_REQUIRE(n, SPC_ERR_INVALID, "spc_get_tree_points: n is null");
    *n = spc::train_tree_points(c, eye_side, max_size, out_host, cap);
    SPC_API_END
}

int spc_tree_to_device(spc_context* ctx, int eye_side, const spc_tree_node* nodes_host, int n, spc_tree_node** dev_out) {
    SPC_API_BEGIN
    SPC_REQUIRE(nodes_host && n > 0 && dev_out, SPC_ERR_INVALID, "spc_tree_to_device: bad arguments");
    spc::tree_install(c, eye_side, nodes_host, n, true);
    *dev_out = (eye_side ? c.train.eye_tree : c.train.light_tree).p;
    SPC_API_END
}
// Device-side counterpart of spc_get_tree_points + spc_build_tree + spc_tree_to_device: the weighted points never leave the device.
int spc_build_tree_from_training_set(spc_context* ctx, int eye_side, int max_size, int subspaces, int label_bias, spc_tree_node** dev_out, int* n_nodes,
                                     spc_tree_node* nodes_host, int cap) {
    SPC_API_BEGIN
    SPC_REQUIRE(dev_out && subspaces >= 1, SPC_ERR_INVALID, "spc_build_tree_from_training_set: bad arguments");
    const int n_pts = spc::train_tree_points_device(c, eye_side, max_size);
    SPC_REQUIRE(n_pts >= 2, SPC_ERR_INVALID, "spc_build_tree_from_training_set: the training set holds %d points", n_pts);
    spc::DevBuf<spc_tree_node>& b = eye_side ? c.train.eye_tree : c.train.light_tree;
    spc::ctree_register(b.p, nullptr, &c);
    const int n = spc::tree_build_device(c, c.train.tree_pts.p, n_pts, subspaces, label_bias, b, nullptr);
    std::vector<spc_tree_node> h((size_t)n);
    SPC_CUDA(cudaMemcpyAsync(h.data(), b.p, (size_t)n * sizeof(spc_tree_node), cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    spc::tree_install(c, eye_side, h.data(), n, false);
    if (n_nodes) *n_nodes = n;
    if (nodes_host && n <= cap) memcpy(nodes_host, h.data(), (size_t)n * sizeof(spc_tree_node));
    *dev_out = b.p;
    SPC_API_END
}
// spc_build_tree on the GPU of `ctx`: same arguments and result (host samples in, host nodes out), for callers that hold the samples
int spc_build_tree_gpu(spc_context* ctx, const spc_divide_weight* samples_host, int n, int K, int label_bias, spc_tree_node* out_host, int cap, int* max_label,
                       int* n_nodes) {
    SPC_API_BEGIN
    SPC_REQUIRE(samples_host && n >= 2 && K >= 1 && n_nodes, SPC_ERR_INVALID, "spc_build_tree_gpu: bad arguments");
    spc::DevBuf<spc_divide_weight> s;
    spc::DevBuf<spc_tree_node> nodes;
    s.alloc(n);
    SPC_CUDA(cudaMemcpyAsync(s.p, samples_host, (size_t)n * sizeof(spc_divide_weight), cudaMemcpyHostToDevice, c.stream));
    const int cnt = spc::tree_build_device(c, s.p, n, K, label_bias, nodes, max_label);
    *n_nodes = cnt;
    if (out_host && cnt <= cap) SPC_CUDA(cudaMemcpyAsync(out_host, nodes.p, (size_t)cnt * sizeof(spc_tree_node), cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    SPC_API_END
}
int spc_preprocess_getQ(spc_context* ctx, const spc_vertex* lvc_dev, const uint8_t* valid_dev, int count_range, int reset, float** Q_dev, int* acc_paths) {
    SPC_API_BEGIN
    SPC_REQUIRE(lvc_dev && valid_dev && count_range > 0, SPC_ERR_INVALID, "spc_preprocess_getQ: bad arguments");
    const int acc = spc::train_get_Q(c, lvc_dev, valid_dev, count_range, reset);
    if (acc_paths) *acc_paths = acc;
    if (Q_dev) *Q_dev = c.train.Q.p;
    SPC_API_END
}
int spc_Q_zero_handle(spc_context* ctx) {
    SPC_API_BEGIN
    spc::train_Q_zero_handle(c);
    SPC_API_END
}
int spc_node_label(spc_context* ctx, const spc_tree_node* eye_tree_dev, const spc_tree_node* light_tree_dev) {
    SPC_API_BEGIN
    spc::train_node_label(c, eye_tree_dev, light_tree_dev);
    SPC_API_END
}
int spc_build_optimal_E_train_data(spc_context* ctx, int n_samples) {
    SPC_API_BEGIN
    spc::train_build_data(c, n_samples);
    SPC_API_END
}
int spc_preprocess_getGamma(spc_context* ctx, float** gamma_dev) {
    SPC_API_BEGIN
    float* g = spc::train_get_gamma(c);
    if (gamma_dev) *gamma_dev = g;
    SPC_API_END
}
int spc_train_optimal_E(spc_context* ctx, int batch_size, int epochs, float lr, float** gamma_dev, float* loss_host, int loss_cap, int* n_batches) {
    SPC_API_BEGIN
    float* g = spc::train_optimal_E(c, batch_size > 0 ? batch_size : 20000, epochs > 0 ? epochs : 1, lr > 0 ? lr : 0.01f, loss_host, loss_cap, n_batches);
    if (gamma_dev) *gamma_dev = g;
    SPC_API_END
}
int spc_Gamma2CMFGamma(spc_context* ctx, const float* gamma_dev, float** cmf_dev) {
    SPC_API_BEGIN
    SPC_REQUIRE(gamma_dev && cmf_dev, SPC_ERR_INVALID, "spc_Gamma2CMFGamma: bad arguments");
    *cmf_dev = spc::train_gamma_to_cmf(c, gamma_dev);
    SPC_API_END
}
int spc_train_set_size(spc_context* ctx, int* n_paths, int* n_conns) {
    SPC_API_BEGIN
    if (n_paths) *n_paths = (int)c.train.n_paths;
    if (n_conns) *n_conns = (int)c.train.n_conns;
    SPC_API_END
}
int spc_train_set_read(spc_context* ctx, spc_train_path* paths_host, spc_train_conn* conns_host) {
    SPC_API_BEGIN
    if (paths_host && c.train.n_paths) SPC_CUDA(cudaMemcpyAsync(paths_host, c.train.paths.p, c.train.n_paths * sizeof(spc_train_path), cudaMemcpyDeviceToHost, c.stream));
    if (conns_host && c.train.n_conns) SPC_CUDA(cudaMemcpyAsync(conns_host, c.train.conns.p, c.train.n_conns * sizeof(spc_train_conn), cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    SPC_API_END
}
int spc_train_data_read(spc_context* ctx, int* N, int* M, float* outlier_threshold, float* f_square, float* pdf0, int* P2N, float* peak, int* label_E, int* label_P) {
    SPC_API_BEGIN
    spc::TrainBuffers& t = c.train;
    if (N) *N = t.N;
    if (M) *M = t.M;
    if (outlier_threshold) *outlier_threshold = t.outlier_threshold;
    auto dl = [&](void* h, const void* d, size_t bytes) {
        if (h && bytes) SPC_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c.stream));
    };
    dl(f_square, t.f_square.p, (size_t)t.N * 4); dl(pdf0, t.pdf0.p, (size_t)t.N * 4); dl(P2N, t.P2N.p, (size_t)t.N * 4);
    dl(peak, t.peak.p, (size_t)t.M * 4); dl(label_E, t.label_E.p, (size_t)t.M * 4); dl(label_P, t.label_P.p, (size_t)t.M * 4);
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    SPC_API_END
}
int spc_train_set_write(spc_context* ctx, const spc_train_path* paths_host, int n_paths, const spc_train_conn* conns_host, int n_conns) {
    SPC_API_BEGIN
    SPC_REQUIRE(n_paths >= 0 && n_conns >= 0 && (paths_host || !n_paths) && (conns_host || !n_conns), SPC_ERR_INVALID, "spc_train_set_write: bad arguments");
    for (int i = 0; i < n_paths; i++)
        SPC_REQUIRE(paths_host[i].begin_ind >= 0 && paths_host[i].begin_ind <= paths_host[i].end_ind && paths_host[i].end_ind <= n_conns, SPC_ERR_INVALID,
                    "spc_train_set_write: path %d spans connections [%d, %d) of %d", i, paths_host[i].begin_ind, paths_host[i].end_ind, n_conns);
    c.train.paths.alloc((size_t)n_paths);
    c.train.conns.alloc((size_t)n_conns);
    if (n_paths) SPC_CUDA(cudaMemcpyAsync(c.train.paths.p, paths_host, (size_t)n_paths * sizeof(spc_train_path), cudaMemcpyHostToDevice, c.stream));
    if (n_conns) SPC_CUDA(cudaMemcpyAsync(c.train.conns.p, conns_host, (size_t)n_conns * sizeof(spc_train_conn), cudaMemcpyHostToDevice, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    c.train.n_paths = (size_t)n_paths;
    c.train.n_conns = (size_t)n_conns;
    c.train.N = c.train.M = 0;
    SPC_API_END
}
int spc_train_Q_write(spc_context* ctx, const float* Q_host, int acc_paths) {
    SPC_API_BEGIN
    SPC_REQUIRE(Q_host && acc_paths >= 0, SPC_ERR_INVALID, "spc_train_Q_write: bad arguments");
    c.train.Q.alloc(c.K);
    SPC_CUDA(cudaMemcpyAsync(c.train.Q.p, Q_host, (size_t)c.K * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    c.train.has_Q = true;
    c.train.acc_valid_path = acc_paths;
    SPC_API_END
}
int spc_train_reset(spc_context* ctx) {
    SPC_API_BEGIN
    c.train.n_paths = c.train.n_conns = 0;
    c.train.has_Q = false;
    c.train.acc_valid_path = 0;
    c.train.N = c.train.M = 0;
    SPC_API_END
}
int spc_device_alloc(spc_context* ctx, size_t bytes, void** dev_out) {
    SPC_API_BEGIN
    SPC_REQUIRE(dev_out, SPC_ERR_INVALID, "spc_device_alloc: dev_out is null");
    SPC_CUDA(cudaMalloc(dev_out, bytes ? bytes : 1));
    SPC_CUDA(cudaMemsetAsync(*dev_out, 0, bytes ? bytes : 1, c.stream));
    SPC_API_END
}
int spc_device_free(spc_context* ctx, void* dev) {
    SPC_API_BEGIN
    if (dev) SPC_CUDA(cudaFree(dev));
    SPC_API_END
}
int spc_upload(spc_context* ctx, void* dev, const void* host, size_t bytes) {
    SPC_API_BEGIN
    SPC_REQUIRE((dev && host) || !bytes, SPC_ERR_INVALID, "spc_upload: null pointer");
    if (bytes) SPC_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    SPC_API_END
}
int spc_download(spc_context* ctx, void* host, const void* dev, size_t bytes) {
    SPC_API_BEGIN
    SPC_REQUIRE((dev && host) || !bytes, SPC_ERR_INVALID, "spc_download: null pointer");
    if (bytes) SPC_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    SPC_API_END
}

}  // extern "C"
