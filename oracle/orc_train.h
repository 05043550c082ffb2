// orc_train.h -- CPU oracle of the subspace-training path (see orc_train.cpp).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <vector>
#include "orc_render.h"

namespace orc {

void pretrace_core(const Frame& fr, int launch_index);   // __raygen__TrainData for one launch index

struct TrainData {   // matrix_parameter::train_data (device_thrust.cu:1600-1611)
    int N = 0, M = 0;
    float outlier_threshold = 0.f;
    std::vector<float> f_square, pdf0, peak;
    std::vector<int> P2N, label_E, label_P;
};

struct TrainSet {   // the file-static neat_paths / neat_conns of device_thrust.cu:428-429
    std::vector<spc_train_path> paths;
    std::vector<spc_train_conn> conns;
    int gather(const spc_train_path* raw_paths, int max_paths, const spc_train_conn* raw_conns, int max_conns);
    void reweight();
    std::vector<spc_divide_weight> tree_points(bool eye_side, int max_size) const;
    void label(const spc_tree_node* eye_tree, const spc_tree_node* light_tree);
    void build_train_data(int n_samples, const float* Q, int K, TrainData& td);
    void gamma_histogram(int K, std::vector<float>& G) const;
};

struct QEstimator {   // preprocess_getQ state (device_thrust.cu:333-334, 349-361)
    int K;
    int acc_valid_path = 0;
    std::vector<float> Q;
    explicit QEstimator(int K_) : K(K_), Q(K_, 0.f) {}
    int add(const spc_vertex* lvc, const uint8_t* valid, int n);
    void zero_handle();
};

void train_gamma(const TrainData& td, int K, std::vector<float>& G, int batch_size, int epochs, float lr, double conservative, std::vector<float>* loss_log);
void gamma_to_cmf(const std::vector<float>& G, int K, float conservative, std::vector<float>& cmf);

}  // namespace orc
