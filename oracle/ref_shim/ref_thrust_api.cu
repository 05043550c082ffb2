// ref_thrust_api.cu -- C entry points around the reference's OWN MyThrustOp library
// (cuda_thrust/device_thrust.cu, compiled by oracle/Makefile from /root/reference with nvcc for sm_100a into
// oracle/_ref/libref_thrust.so).  TEST INFRASTRUCTURE: it is the GPU-side checker of the post-processing seam
// (LVC_Process, training-set plumbing, Q, Gamma, the Adam trainer); it needs a CUDA device, so only the
// `-m gpu` tests load it.  NUM_SUBSPACE is compiled in as 1000 (optixPathTracer.h:31).
//
// The reference keeps all state in file-static thrust vectors (device_thrust.cu:287-293,428-429,...): one
// training set per process, exactly like the reference application.
#include <cstring>
#include <vector>
#include "device_thrust.h"

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API int ref_thrust_num_subspace() { return NUM_SUBSPACE; }

REF_API void ref_thrust_lvc_process(void* lvc_dev, void* valid_dev, int n, SubspaceSampler* out) {
    *out = MyThrustOp::LVC_Process(thrust::device_ptr<BDPTVertex>((BDPTVertex*)lvc_dev), thrust::device_ptr<bool>((bool*)valid_dev), n);
}
REF_API int ref_thrust_valid_sample_gather(void* paths_dev, int n_paths, void* conns_dev, int n_conns) {
    return MyThrustOp::valid_sample_gather(thrust::device_ptr<preTracePath>((preTracePath*)paths_dev), n_paths,
                                           thrust::device_ptr<preTraceConnection>((preTraceConnection*)conns_dev), n_conns);
}
REF_API void ref_thrust_sample_reweight() { MyThrustOp::sample_reweight(); }
REF_API int ref_thrust_get_tree_points(int eye_side, int max_size, void* out, int cap) {
    std::vector<classTree::divide_weight> v = MyThrustOp::get_weighted_point_for_tree_building(eye_side != 0, max_size);
    const int n = (int)v.size();
    if (n <= cap) memcpy(out, v.data(), n * sizeof(classTree::divide_weight));
    return n;
}
REF_API void* ref_thrust_tree_to_device(int eye_side, void* nodes_host, int n) {
    return eye_side ? (void*)MyThrustOp::eye_tree_to_device((classTree::tree_node*)nodes_host, n)
                    : (void*)MyThrustOp::light_tree_to_device((classTree::tree_node*)nodes_host, n);
}
static thrust::device_ptr<float> g_Q, g_Gamma;
REF_API int ref_thrust_get_Q(void* lvc_dev, void* valid_dev, int n, int reset) {
    if (reset) g_Q = thrust::device_ptr<float>();
    return MyThrustOp::preprocess_getQ(thrust::device_ptr<BDPTVertex>((BDPTVertex*)lvc_dev), thrust::device_ptr<bool>((bool*)valid_dev), n, g_Q);
}
REF_API void ref_thrust_Q_zero_handle() { MyThrustOp::Q_zero_handle(g_Q); }
REF_API void* ref_thrust_Q_ptr() { return thrust::raw_pointer_cast(g_Q); }
REF_API void ref_thrust_node_label(void* eye_tree_dev, void* light_tree_dev) {
    MyThrustOp::node_label((classTree::tree_node*)eye_tree_dev, (classTree::tree_node*)light_tree_dev);
}
REF_API void ref_thrust_build_train_data(int n_samples) { MyThrustOp::build_optimal_E_train_data(n_samples); }
REF_API void* ref_thrust_get_gamma() {
    MyThrustOp::preprocess_getGamma(g_Gamma);
    return thrust::raw_pointer_cast(g_Gamma);
}
REF_API void* ref_thrust_train_gamma() {
    MyThrustOp::train_optimal_E(g_Gamma);
    return thrust::raw_pointer_cast(g_Gamma);
}
REF_API void* ref_thrust_gamma_to_cmf() { return thrust::raw_pointer_cast(MyThrustOp::Gamma2CMFGamma(g_Gamma)); }
REF_API int ref_thrust_download(const void* dev, void* host, size_t bytes) {
    return (int)cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost);
}

// the reference's file-scope state (device_thrust.cu:428-429, 3098-3111), read back for comparison
namespace MyThrustOp {
extern thrust::device_vector<preTracePath> neat_paths;
extern thrust::device_vector<preTraceConnection> neat_conns;
extern thrust::device_vector<float> b_f_square, b_pdf0, b_pdf_peak;
extern thrust::device_vector<int> b_label_E, b_label_P, b_P2N_ind_d;
}
REF_API void ref_thrust_set_sizes(int* n_paths, int* n_conns) {
    *n_paths = (int)MyThrustOp::neat_paths.size();
    *n_conns = (int)MyThrustOp::neat_conns.size();
}
REF_API void ref_thrust_set_read(void* paths_host, void* conns_host) {
    cudaMemcpy(paths_host, thrust::raw_pointer_cast(MyThrustOp::neat_paths.data()), MyThrustOp::neat_paths.size() * sizeof(preTracePath), cudaMemcpyDeviceToHost);
    cudaMemcpy(conns_host, thrust::raw_pointer_cast(MyThrustOp::neat_conns.data()), MyThrustOp::neat_conns.size() * sizeof(preTraceConnection), cudaMemcpyDeviceToHost);
}
REF_API void ref_thrust_train_data_sizes(int* N, int* M) {
    *N = (int)MyThrustOp::b_f_square.size();
    *M = (int)MyThrustOp::b_pdf_peak.size();
}
REF_API void ref_thrust_train_data_read(float* f_square, float* pdf0, int* P2N, float* peak, int* label_E, int* label_P) {
    using namespace MyThrustOp;
    cudaMemcpy(f_square, thrust::raw_pointer_cast(b_f_square.data()), b_f_square.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(pdf0, thrust::raw_pointer_cast(b_pdf0.data()), b_pdf0.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(P2N, thrust::raw_pointer_cast(b_P2N_ind_d.data()), b_P2N_ind_d.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(peak, thrust::raw_pointer_cast(b_pdf_peak.data()), b_pdf_peak.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(label_E, thrust::raw_pointer_cast(b_label_E.data()), b_label_E.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(label_P, thrust::raw_pointer_cast(b_label_P.data()), b_label_P.size() * 4, cudaMemcpyDeviceToHost);
}
