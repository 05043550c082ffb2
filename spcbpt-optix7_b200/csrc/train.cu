// train.cu -- the subspace-training plumbing behind the MyThrustOp seam (cuda_thrust/device_thrust.cu), on the
// device.  The reference does most of these steps in serial host loops after copying the whole training set
// (~600 MB of structs) to the host and back; here the set never leaves HBM.
//
//   valid_sample_gather         (:457-493)   flag scan + order-preserving scatter + index fix-up, appended to the set
//   sample_reweight             (:574-623)   ordered binning on the 10x10-pixel grid, then a per-path division
//   get_weighted_point_...      (:494-527)   one lane per connection
//   node_label                  (:554-573)   classification of both endpoints of every connection
//   preprocess_getQ / Q_zero    (:347-409, :335-346)  ordered binning of the LVC (same order as the host loop)
//   build_optimal_E_train_data  (:3261-3325) outlier clamp + SoA training arrays
//   preprocess_getGamma         (:627-667)   K x K scatter-add histogram (fp32 atomics) + row normalise
//   train_optimal_E             (:3327-3344 -> matrix_parameter::fit :1615-1655, matrix_optimal_operator :923-1190,
//                                adam_step_func :1437-1470): three fused kernels per batch instead of ~25 thrust calls
//   Gamma2CMFGamma              (:3406-3433) one lane per row, sequential fp32 prefix (the reference's order)
// Not a dense contraction anywhere: CUDA cores + fp32 atomics, no tensor cores.
#include <algorithm>
#include <cfloat>
#include <cub/cub.cuh>
#include <iterator>
#include <map>
#include <mutex>
#include "shade.cuh"

namespace spc {

// ---------------------------------------------------------------------------------------------
// small device-wide exclusive scan of int flags (three kernels; n up to 2^31)
// ---------------------------------------------------------------------------------------------
constexpr int kScanBlock = 256, kScanPer = 8, kScanTile = kScanBlock * kScanPer;

__global__ void k_scan_local(const int* __restrict__ in, int* __restrict__ out, int n, int* __restrict__ sums) {
    __shared__ int s_warp[kScanBlock / 32];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    int v[kScanPer], tot = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        tot += v[k];
    }
    int inc = tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < kScanBlock / 32 ? s_warp[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < kScanBlock / 32) s_warp[lane] = w;
    }
    __syncthreads();
    int run = inc - tot + (warp ? s_warp[warp - 1] : 0);
#pragma unroll
    for (int k = 0; k < kScanPer; k++) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == kScanBlock - 1) sums[blockIdx.x] = s_warp[kScanBlock / 32 - 1];
}
__global__ void k_scan_sums(int* __restrict__ sums, int nb, int* __restrict__ total) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < nb; i++) { const int v = sums[i]; sums[i] = run; run += v; }
        *total = run;
    }
}
__global__ void k_scan_add(int* __restrict__ out, int n, const int* __restrict__ sums) {
    const int i = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    const int add = sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanPer; k++)
        if (i + k < n) out[i + k] += add;
}
static void exclusive_scan(Context& c, const int* in, int* out, int n, DevBuf<int>& sums, int* total_dev) {
    const int nb = (n + kScanTile - 1) / kScanTile;
    sums.alloc(nb + 1);
    k_scan_local<<<nb, kScanBlock, 0, c.stream>>>(in, out, n, sums.p);
    k_scan_sums<<<1, 32, 0, c.stream>>>(sums.p, nb, total_dev);
    k_scan_add<<<nb, kScanBlock, 0, c.stream>>>(out, n, sums.p);
    c.launches += 3;
}

// ---------------------------------------------------------------------------------------------
// valid_sample_gather
// ---------------------------------------------------------------------------------------------
__global__ void k_path_flags(const spc_train_path* __restrict__ p, int n, int* __restrict__ f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = p[i].valid ? 1 : 0;
}
__global__ void k_conn_flags(const spc_train_conn* __restrict__ p, int n, int* __restrict__ f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = p[i].valid ? 1 : 0;
}
__global__ void k_gather_conns(const spc_train_conn* __restrict__ raw, int n, const int* __restrict__ pos, spc_train_conn* __restrict__ neat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && raw[i].valid) neat[pos[i]] = raw[i];
}
// bias_arrange_op (device_thrust.cu:431-452) fused with the path copy
__global__ void k_gather_paths(const spc_train_path* __restrict__ raw, int n, const int* __restrict__ ppos, const int* __restrict__ cpos,
                               spc_train_path* __restrict__ neat_paths, spc_train_conn* __restrict__ neat_conns, int sample_bias, int node_bias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !raw[i].valid) return;
    spc_train_path s = raw[i];
    const int id = ppos[i];
    const int len = s.end_ind - s.begin_ind;
    const int b = cpos[s.begin_ind];
    s.begin_ind = node_bias + b;
    s.end_ind = node_bias + b + len;
    neat_paths[sample_bias + id] = s;
    for (int k = 0; k < len; k++) neat_conns[node_bias + b + k].path_id = id + sample_bias;
}

template <typename T>
static void grow_keep(DevBuf<T>& buf, size_t used, size_t need, cudaStream_t st, size_t reserve = 0) {
    if (need <= buf.n && buf.p) return;
    size_t cap = std::max<size_t>(std::max(need, reserve), buf.n * 2);
    T* np = nullptr;
    SPC_CUDA(cudaMalloc((void**)&np, cap * sizeof(T)));
    if (used && buf.p) SPC_CUDA(cudaMemcpyAsync(np, buf.p, used * sizeof(T), cudaMemcpyDeviceToDevice, st));
    SPC_CUDA(cudaStreamSynchronize(st));
    if (buf.p) cudaFree(buf.p);
    buf.p = np;
    buf.n = cap;
}

static void pinned(Context& c) {
    if (!c.h_pinned) SPC_CUDA(cudaMallocHost((void**)&c.h_pinned, 64 * sizeof(int)));
}

int train_gather(Context& c, const spc_train_path* raw_paths, int max_paths, const spc_train_conn* raw_conns, int max_conns) {
    TrainBuffers& t = c.train;
    pinned(c);
    cudaStream_t st = c.stream;
    t.flag_p.alloc(max_paths); t.flag_c.alloc(max_conns); t.pos_p.alloc(max_paths); t.pos_c.alloc(max_conns); t.totals.alloc(8);
    k_path_flags<<<(max_paths + 255) / 256, 256, 0, st>>>(raw_paths, max_paths, t.flag_p.p);
    k_conn_flags<<<(max_conns + 255) / 256, 256, 0, st>>>(raw_conns, max_conns, t.flag_c.p);
    c.launches += 2;
    exclusive_scan(c, t.flag_p.p, t.pos_p.p, max_paths, t.scan_sums, t.totals.p);
    exclusive_scan(c, t.flag_c.p, t.pos_c.p, max_conns, t.scan_sums2, t.totals.p + 1);
    SPC_CUDA(cudaMemcpyAsync(c.h_pinned, t.totals.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SPC_CUDA(cudaStreamSynchronize(st));
    const int sample_count = c.h_pinned[0], node_count = c.h_pinned[1];
    // option "train_reserve_paths": the caller's target size of the training set.  Both arrays are then sized once, at the first
    // gather (connections per path extrapolated from this launch, +15 %), instead of doubling ~8 times up to ~0.6 GB with a
    // cudaMalloc + copy + cudaFree each time (0.1 -> 0.4 s of a 2 M-path schedule on a cold allocator).
    size_t reserve_p = 0, reserve_c = 0;
    const size_t hint = (size_t)c.opt[OPT_TRAIN_RESERVE];
    if (hint > t.n_paths + (size_t)sample_count && sample_count > 0) {
        const double per_path = (double)(t.n_conns + (size_t)node_count) / (double)(t.n_paths + (size_t)sample_count);
        reserve_p = hint + (size_t)max_paths;
        reserve_c = (size_t)((double)reserve_p * per_path * 1.15) + (size_t)max_conns;
    }
    grow_keep(t.paths, t.n_paths, t.n_paths + sample_count, st, reserve_p);
    grow_keep(t.conns, t.n_conns, t.n_conns + node_count, st, reserve_c);
    k_gather_conns<<<(max_conns + 255) / 256, 256, 0, st>>>(raw_conns, max_conns, t.pos_c.p, t.conns.p + t.n_conns);
    k_gather_paths<<<(max_paths + 255) / 256, 256, 0, st>>>(raw_paths, max_paths, t.pos_p.p, t.pos_c.p, t.paths.p, t.conns.p, (int)t.n_paths, (int)t.n_conns);
    c.launches += 2;
    SPC_CUDA(cudaGetLastError());
    t.n_paths += sample_count;
    t.n_conns += node_count;
    return sample_count;
}

// ---------------------------------------------------------------------------------------------
// sample_reweight: weight[block] = sum of contri/pdf over the paths of a 10x10-pixel block IN PATH ORDER
// (grid stride hard-coded to 192 = 1920/10, table of 1920*1000/100*1.1 entries, as in the reference)
// ---------------------------------------------------------------------------------------------
constexpr int kReweightBins = (int)(1920 * 1000 / 100 * 1.1);
__global__ void k_reweight_keys(const spc_train_path* __restrict__ p, int n, int* __restrict__ key, float* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int n_id = p[i].pixel_x / 10 + p[i].pixel_y / 10 * 192;
    const float ww = (p[i].contri.x + p[i].contri.y + p[i].contri.z) / p[i].sample_pdf;
    const bool skip = isnan(ww) || isinf(ww) || n_id < 0 || n_id >= kReweightBins;
    key[i] = skip ? -1 : n_id;
    w[i] = skip ? 0.f : ww;
}
__global__ void k_reweight_apply(spc_train_path* __restrict__ p, int n, const spc_subspace* __restrict__ bins) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int n_id = p[i].pixel_x / 10 + p[i].pixel_y / 10 * 192;
    const float sum = (n_id >= 0 && n_id < kReweightBins) ? bins[n_id].sum_pmf : 0.f;
    const float w = (float)((double)(sum / 100) + 0.1);
    const float3 c = f3(p[i].contri.x, p[i].contri.y, p[i].contri.z) / w;
    p[i].contri = spc_float3{c.x, c.y, c.z};
}
// sum_pmf of every bin <-> a dense float array (for the all-reduce of a sharded training set)
__global__ void k_bins_get(const spc_subspace* __restrict__ bins, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = bins[i].sum_pmf;
}
__global__ void k_bins_put(spc_subspace* __restrict__ bins, int n, const float* __restrict__ in) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bins[i].sum_pmf = in[i];
}
void train_reweight(Context& c) {
    TrainBuffers& t = c.train;
    const int n = (int)t.n_paths;
    if (n == 0 && c.comm_world <= 1) return;
    LvcBuffers& b = c.bins_tmp;
    b.key.alloc(n); b.weight.alloc(n); b.totals.alloc(kReweightBins + 8);
    SPC_CUDA(cudaMemsetAsync(b.totals.p + kReweightBins, 0, 8 * sizeof(int), c.stream));
    if (n) k_reweight_keys<<<(n + 255) / 256, 256, 0, c.stream>>>(t.paths.p, n, b.key.p, b.weight.p);
    bin_ordered(c, b, n, kReweightBins, b.totals.p + kReweightBins);
    if (c.comm_world > 1) {   // the grid sums run over the whole training set: add the other ranks' shards
        t.sort_vals.alloc(kReweightBins);
        k_bins_get<<<(kReweightBins + 255) / 256, 256, 0, c.stream>>>(b.subspace.p, kReweightBins, t.sort_vals.p);
        comm_allreduce_sum(c, t.sort_vals.p, kReweightBins);
        k_bins_put<<<(kReweightBins + 255) / 256, 256, 0, c.stream>>>(b.subspace.p, kReweightBins, t.sort_vals.p);
        c.launches += 2;
    }
    if (n) k_reweight_apply<<<(n + 255) / 256, 256, 0, c.stream>>>(t.paths.p, n, b.subspace.p);
    c.launches += 2;
    SPC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// get_weighted_point_for_tree_building: connections of the first `limit` paths are the first paths[limit-1].end_ind
// entries of the connection array (paths and their connections are both stored in path order)
// ---------------------------------------------------------------------------------------------
__global__ void k_tree_points(const spc_train_path* __restrict__ paths, const spc_train_conn* __restrict__ conns, int n_conn, int eye_side,
                              spc_divide_weight* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_conn) return;
    const spc_train_conn& s = conns[j];
    spc_divide_weight t;
    t.position = t.dir = t.normal = spc_float3{0.f, 0.f, 0.f};
    t.weight = 0.f;   // the reference pushes an uninitialised record for emitter endpoints on the light side
    const spc_train_path& p = paths[s.path_id];
    float w = (p.contri.x + p.contri.y + p.contri.z) / p.sample_pdf;
    // deviation: a NaN/Inf weight (sample_pdf underflowed to 0 on a long path) would poison every sum of the
    // tree builder -- the reference then degenerates to a single subspace; such samples get weight 0 here
    if (isnan(w) || isinf(w)) w = 0.f;
    if (eye_side) {
        t.dir = s.A_dir; t.normal = s.A_normal; t.position = s.A_position; t.weight = w;
    } else if (!s.light_source) {
        t.dir = s.B_dir; t.normal = s.B_normal; t.position = s.B_position; t.weight = w;
    }
    out[j] = t;
}
// the weighted points stay on the device (t.tree_pts): input of the device-side tree build (tree_build.cu)
int train_tree_points_device(Context& c, int eye_side, int max_size) {
    TrainBuffers& t = c.train;
    if (t.n_paths == 0) return 0;
    pinned(c);
    const size_t limit = max_size == 0 ? t.n_paths : std::min(t.n_paths, (size_t)max_size);
    spc_train_path last;
    SPC_CUDA(cudaMemcpyAsync(&last, t.paths.p + limit - 1, sizeof(last), cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    const int n = last.end_ind;
    t.tree_pts.alloc(n);
    if (n) k_tree_points<<<(n + 255) / 256, 256, 0, c.stream>>>(t.paths.p, t.conns.p, n, eye_side, t.tree_pts.p);
    c.launches++;
    SPC_CUDA(cudaGetLastError());
    return n;
}
int train_tree_points(Context& c, int eye_side, int max_size, spc_divide_weight* out_host, int cap) {
    TrainBuffers& t = c.train;
    if (t.n_paths == 0) return 0;
    if (!out_host) {   // size query
        pinned(c);
        const size_t limit = max_size == 0 ? t.n_paths : std::min(t.n_paths, (size_t)max_size);
        spc_train_path last;
        SPC_CUDA(cudaMemcpyAsync(&last, t.paths.p + limit - 1, sizeof(last), cudaMemcpyDeviceToHost, c.stream));
        SPC_CUDA(cudaStreamSynchronize(c.stream));
        return last.end_ind;
    }
    const int n = train_tree_points_device(c, eye_side, max_size);
    if (n > cap) return n;
    SPC_CUDA(cudaMemcpyAsync(out_host, t.tree_pts.p, (size_t)n * sizeof(spc_divide_weight), cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    return n;
}

// ---------------------------------------------------------------------------------------------
// node_label
// ---------------------------------------------------------------------------------------------
__global__ void k_node_label(spc_train_conn* __restrict__ conns, int n, const spc_tree_node* __restrict__ eye_tree, const spc_tree_node* __restrict__ light_tree) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    spc_train_conn& s = conns[i];
    s.label_A = tree_label(eye_tree, f3(s.A_position.x, s.A_position.y, s.A_position.z), f3(s.A_normal.x, s.A_normal.y, s.A_normal.z));
    if (!s.light_source) s.label_B = tree_label(light_tree, f3(s.B_position.x, s.B_position.y, s.B_position.z), f3(s.B_normal.x, s.B_normal.y, s.B_normal.z));
}
void train_node_label(Context& c, const spc_tree_node* eye_tree, const spc_tree_node* light_tree) {
    TrainBuffers& t = c.train;
    const int n = (int)t.n_conns;
    if (!n) return;
    k_node_label<<<(n + 255) / 256, 256, 0, c.stream>>>(t.conns.p, n, eye_tree, light_tree);
    c.launches++;
    SPC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// preprocess_getQ: Q = running mean (weighted by path counts) of  sum_{v in subspace} w_v / path_count
// ---------------------------------------------------------------------------------------------
__global__ void k_q_update(float* __restrict__ Q, const spc_subspace* __restrict__ sub, int K, int path_count, float t) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const float tmp = sub[i].sum_pmf / (float)path_count;
    Q[i] = Q[i] * (1 - t) + tmp * t;
}
__global__ void k_q_zero_handle(float* __restrict__ Q, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K && Q[i] == 0) Q[i] = FLT_MAX;
}
int train_get_Q(Context& c, const spc_vertex* lvc, const uint8_t* valid, int n, int reset) {
    TrainBuffers& t = c.train;
    pinned(c);
    const int K = c.K;
    if (reset || !t.has_Q) {
        t.Q.alloc(K);
        SPC_CUDA(cudaMemsetAsync(t.Q.p, 0, K * sizeof(float), c.stream));
        t.acc_valid_path = 0;
        t.has_Q = true;
    }
    int* counters = lvc_bin(c, c.bins_tmp, lvc, valid, n);
    SPC_CUDA(cudaMemcpyAsync(c.h_pinned, counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    const int path_count = c.h_pinned[1];
    t.acc_valid_path += path_count;
    const float tt = (float)path_count / (float)t.acc_valid_path;
    k_q_update<<<(K + 127) / 128, 128, 0, c.stream>>>(t.Q.p, c.bins_tmp.subspace.p, K, path_count, tt);
    c.launches++;
    SPC_CUDA(cudaGetLastError());
    return t.acc_valid_path;
}
__global__ void k_q_scale(float* __restrict__ Q, int K, float s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) Q[i] *= s;
}
// spc_allreduce_training_stats: Q <- sum_r n_r Q_r / sum_r n_r (n_r = light paths behind rank r's estimate)
void train_allreduce_Q(Context& c) {
    TrainBuffers& t = c.train;
    SPC_REQUIRE(t.has_Q, SPC_ERR_INVALID, "spc_allreduce_training_stats before preprocess_getQ");
    if (c.comm_world <= 1) return;
    const int K = c.K;
    t.sort_vals.alloc(K + 1);
    k_q_scale<<<(K + 127) / 128, 128, 0, c.stream>>>(t.Q.p, K, (float)t.acc_valid_path);
    const float n_r = (float)t.acc_valid_path;
    SPC_CUDA(cudaMemcpyAsync(t.sort_vals.p, t.Q.p, K * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
    SPC_CUDA(cudaMemcpyAsync(t.sort_vals.p + K, &n_r, sizeof(float), cudaMemcpyHostToDevice, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));   // n_r is a stack variable
    comm_allreduce_sum(c, t.sort_vals.p, (size_t)K + 1);
    float total = 0.f;
    SPC_CUDA(cudaMemcpyAsync(&total, t.sort_vals.p + K, sizeof(float), cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    SPC_REQUIRE(total > 0.f, SPC_ERR_INVALID, "spc_allreduce_training_stats: no light paths on any rank");
    SPC_CUDA(cudaMemcpyAsync(t.Q.p, t.sort_vals.p, K * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
    k_q_scale<<<(K + 127) / 128, 128, 0, c.stream>>>(t.Q.p, K, 1.0f / total);
    t.acc_valid_path = (int)total;
    c.launches += 2;
    SPC_CUDA(cudaGetLastError());
}
void train_Q_zero_handle(Context& c) {
    TrainBuffers& t = c.train;
    SPC_REQUIRE(t.has_Q, SPC_ERR_INVALID, "Q_zero_handle before preprocess_getQ");
    k_q_zero_handle<<<(c.K + 127) / 128, 128, 0, c.stream>>>(t.Q.p, c.K);
    c.launches++;
}

// ---------------------------------------------------------------------------------------------
// build_optimal_E_train_data
// ---------------------------------------------------------------------------------------------
#define SPC_LOSS_THRESHOLD 1000000.0f   // optimal_E_loss_threshold (device_thrust.cu:3097)
__device__ __forceinline__ float outlier_value(const spc_train_path& s, const spc_train_conn* __restrict__ nodes, const float* __restrict__ Q) {
    float outler_value = s.fix_pdf;
    const float weight = s.contri.x + s.contri.y + s.contri.z;
    float loss = weight * weight / s.sample_pdf;
    if (loss > SPC_LOSS_THRESHOLD || isnan(loss)) loss = SPC_LOSS_THRESHOLD;
    for (int i = s.begin_ind; i < s.end_ind; i++)
        outler_value = (float)((double)outler_value + (double)(nodes[i].peak_pdf / Q[nodes[i].label_B]) / 1000.0);   // float += double
    return loss / outler_value;
}
__global__ void k_outlier_values(const spc_train_path* __restrict__ paths, int n, const spc_train_conn* __restrict__ nodes, const float* __restrict__ Q, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = outlier_value(paths[i], nodes, Q);
}
__global__ void k_outlier_clean(spc_train_path* __restrict__ paths, int n, const spc_train_conn* __restrict__ nodes, const float* __restrict__ Q, float threshold) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (outlier_value(paths[i], nodes, Q) > threshold) {
        paths[i].contri.x *= 0; paths[i].contri.y *= 0; paths[i].contri.z *= 0;
    }
}
__global__ void k_td_samples(const spc_train_path* __restrict__ paths, int n, float* __restrict__ f_square, float* __restrict__ pdf0, int* __restrict__ P2N) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    const spc_train_path& s = paths[id];
    const float weight = s.contri.x + s.contri.y + s.contri.z;
    float f = weight * weight / s.sample_pdf;
    if (f > SPC_LOSS_THRESHOLD || isnan(f)) f = SPC_LOSS_THRESHOLD;
    f_square[id] = f;
    pdf0[id] = s.fix_pdf;
    P2N[id] = s.begin_ind;
}
__global__ void k_td_nodes(const spc_train_conn* __restrict__ nodes, int m, const float* __restrict__ Q, int K, float* __restrict__ peak, int* __restrict__ label_E,
                           int* __restrict__ label_P) {
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= m) return;
    const spc_train_conn& s = nodes[id];
    label_E[id] = s.label_A * K + s.label_B;
    label_P[id] = s.path_id;
    float p = (double)Q[s.label_B] > 0.0 ? s.peak_pdf / Q[s.label_B] : 0.0f;
    if (isnan(p) || isinf(p)) p = 0;
    peak[id] = p;
}
void train_build_data(Context& c, int n_samples) {
    TrainBuffers& t = c.train;
    SPC_REQUIRE(t.has_Q, SPC_ERR_INVALID, "build_optimal_E_train_data before preprocess_getQ");
    SPC_REQUIRE(n_samples >= 1000 && (size_t)n_samples <= t.n_paths, SPC_ERR_INVALID, "build_optimal_E_train_data: N=%d but the set holds %zu paths (>= 1000 needed)", n_samples, t.n_paths);
    pinned(c);
    cudaStream_t st = c.stream;
    spc_train_path last;
    SPC_CUDA(cudaMemcpyAsync(&last, t.paths.p + n_samples - 1, sizeof(last), cudaMemcpyDeviceToHost, st));
    t.outlier.alloc(1000);
    k_outlier_values<<<4, 256, 0, st>>>(t.paths.p, 1000, t.conns.p, t.Q.p, t.outlier.p);
    float h[1000];
    SPC_CUDA(cudaMemcpyAsync(h, t.outlier.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    SPC_CUDA(cudaStreamSynchronize(st));
    std::sort(h, h + 1000);   // thrust::sort + [999] (device_thrust.cu:3282-3284); NaNs cannot occur past the loss clamp unless Q is NaN
    t.outlier_threshold = h[999];
    if (c.comm_world > 1) {   // sharded set: the first 1000 paths of the whole set are rank 0's
        SPC_CUDA(cudaMemcpyAsync(t.outlier.p, &t.outlier_threshold, sizeof(float), cudaMemcpyHostToDevice, st));
        comm_bcast(c, t.outlier.p, sizeof(float), 0);
        SPC_CUDA(cudaMemcpyAsync(&t.outlier_threshold, t.outlier.p, sizeof(float), cudaMemcpyDeviceToHost, st));
        SPC_CUDA(cudaStreamSynchronize(st));
    }
    const int np = (int)t.n_paths;
    k_outlier_clean<<<(np + 255) / 256, 256, 0, st>>>(t.paths.p, np, t.conns.p, t.Q.p, t.outlier_threshold);
    t.N = n_samples;
    t.M = last.end_ind;
    t.f_square.alloc(t.N); t.pdf0.alloc(t.N); t.P2N.alloc(t.N); t.peak.alloc(t.M); t.label_E.alloc(t.M); t.label_P.alloc(t.M);
    k_td_samples<<<(t.N + 255) / 256, 256, 0, st>>>(t.paths.p, t.N, t.f_square.p, t.pdf0.p, t.P2N.p);
    k_td_nodes<<<(t.M + 255) / 256, 256, 0, st>>>(t.conns.p, t.M, t.Q.p, c.K, t.peak.p, t.label_E.p, t.label_P.p);
    c.launches += 4;
    SPC_CUDA(cudaGetLastError());
    t.h_P2N.resize(t.N);
    SPC_CUDA(cudaMemcpyAsync(t.h_P2N.data(), t.P2N.p, (size_t)t.N * sizeof(int), cudaMemcpyDeviceToHost, st));
    SPC_CUDA(cudaStreamSynchronize(st));
}

// ---------------------------------------------------------------------------------------------
// preprocess_getGamma
// ---------------------------------------------------------------------------------------------
// The reference adds the clamped path weight of every split to h_Gamma[eye * K + light] in one serial host loop over the
// connections in index order (device_thrust.cu:633-646).  To get the same fp32 sums on the device the (cell, weight) pairs are
// sorted by cell with a stable radix sort and each cell is then summed in order by the thread that owns its first element:
// bit-identical to the serial loop, and run-to-run reproducible (no floating-point atomics).
__global__ void k_gamma_pairs(const spc_train_path* __restrict__ paths, const spc_train_conn* __restrict__ conns, int n_conns, int K,
                              uint32_t* __restrict__ keys, float* __restrict__ vals) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_conns) return;
    const spc_train_conn& s = conns[j];
    const spc_train_path& p = paths[s.path_id];
    const float weight = (p.contri.x + p.contri.y + p.contri.z) / p.sample_pdf;
    keys[j] = (uint32_t)s.label_A * (uint32_t)K + (uint32_t)s.label_B;
    vals[j] = (float)fmin((double)weight, 10.0);
}
// out[key] = in-order sum of the values of the key's run (keys sorted); runs start where the key changes
__global__ void k_segment_sum_ordered(const uint32_t* __restrict__ keys, const float* __restrict__ vals, int n, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t key = keys[j];
    if (j > 0 && keys[j - 1] == key) return;
    // end of the run by bisection, so that the summation loop has a known trip count (loads run ahead of the FADD chain: a
    // hot cell holds 10^4..10^5 entries)
    int lo = j, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] == key) lo = mid;
        else hi = mid;
    }
    float sum = 0.f;
#pragma unroll 8
    for (int k = j; k < hi; k++) sum += vals[k];
    out[key] = sum;
}
static int bits_for(size_t n) {
    int b = 1;
    while (((size_t)1 << b) < n && b < 32) b++;
    return b;
}
__global__ void k_gamma_rownorm(float* __restrict__ G, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    float* row = G + (size_t)i * K;
    float weightS = 0;
    for (int j = 0; j < K; j++) weightS += row[j];
    for (int j = 0; j < K; j++) {
        row[j] /= weightS;
        if (weightS <= 1e-10f) row[j] = (float)(1.0 / K);
    }
}
float* train_get_gamma(Context& c) {
    TrainBuffers& t = c.train;
    const int K = c.K;
    t.gamma.alloc((size_t)K * K);
    SPC_CUDA(cudaMemsetAsync(t.gamma.p, 0, (size_t)K * K * sizeof(float), c.stream));
    const int n = (int)t.n_conns;
    if (n) {
        t.sort_keys.alloc(n); t.sort_keys2.alloc(n); t.sort_vals.alloc(n); t.sort_vals2.alloc(n);
        k_gamma_pairs<<<(n + 255) / 256, 256, 0, c.stream>>>(t.paths.p, t.conns.p, n, K, t.sort_keys.p, t.sort_vals.p);
        size_t tmp_bytes = 0;
        const int end_bit = bits_for((size_t)K * K);
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, t.sort_keys.p, t.sort_keys2.p, t.sort_vals.p, t.sort_vals2.p, n, 0, end_bit, c.stream);
        t.sort_tmp.alloc(tmp_bytes);
        SPC_CUDA(cub::DeviceRadixSort::SortPairs(t.sort_tmp.p, tmp_bytes, t.sort_keys.p, t.sort_keys2.p, t.sort_vals.p, t.sort_vals2.p, n, 0, end_bit, c.stream));
        k_segment_sum_ordered<<<(n + 255) / 256, 256, 0, c.stream>>>(t.sort_keys2.p, t.sort_vals2.p, n, t.gamma.p);
        c.launches += 3;
    }
    comm_allreduce_sum(c, t.gamma.p, (size_t)K * K);   // sharded set: the histogram runs over every rank's connections
    k_gamma_rownorm<<<(K + 63) / 64, 64, 0, c.stream>>>(t.gamma.p, K);
    c.launches += 1;
    SPC_CUDA(cudaGetLastError());
    return t.gamma.p;
}

// ---------------------------------------------------------------------------------------------
// train_optimal_E: minimise sum_i f2_i / (pdf0_i + sum_{n in i} peak_n * E[label_n]) over theta,
// E = (1-c) * sigmoid(theta) / rowsum + c / K, Adam(lr .01, .9, .999, 1e-8); 1 epoch of N/20000 batches
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_ref(float a) { return (float)(1.0 / (1.0 + (double)cm_expf(-a))); }   // sigmoid<float> (:705-712)

// one block per row: E = sigmoid(theta); row sum; E = E / sum * (1-c) + c/K   (get_E, :1149-1172)
__global__ void k_train_E(const float* __restrict__ theta, int K, float c_keep, float c_uniform, float* __restrict__ E, float* __restrict__ Esum, float* __restrict__ dE) {
    __shared__ float s_red[32];
    const int row = blockIdx.x;
    float part = 0.f;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        const float s = sigmoidf_ref(theta[(size_t)row * K + j]);
        E[(size_t)row * K + j] = s;
        dE[(size_t)row * K + j] = 0.f;
        part += s;
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) { s_red[0] = v; Esum[row] = v; }
    }
    __syncthreads();
    const float sum = s_red[0];
    for (int j = threadIdx.x; j < K; j += blockDim.x) E[(size_t)row * K + j] = E[(size_t)row * K + j] / sum * c_keep + c_uniform;
}
// one lane per path of the batch: forward pdf, d(loss)/d(pdf), scatter-add into dE (get_forward_pdfs .. get_dE, :981-1088)
__global__ void k_train_paths(const float* __restrict__ E, const float* __restrict__ f_square, const float* __restrict__ pdf0, const float* __restrict__ peak,
                              const int* __restrict__ label_E, const int* __restrict__ P2N, int bias_sample, int batch, int N, int M,
                              float* __restrict__ path_d, float* __restrict__ path_loss) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) {
        const int b = P2N[bias_sample + i];
        // a path's nodes are [P2N[i], P2N[i+1]).  (The reference hands the last batch every remaining node, device_thrust.cu:1636,
        // which only is well-defined when N is a multiple of the batch size -- its own use, 2 000 000 / 20 000.)
        const int e = (bias_sample + i + 1 < N) ? P2N[bias_sample + i + 1] : M;
        float pdf = 0.f;
        for (int k = b; k < e; k++) pdf += peak[k] * E[label_E[k]];
        pdf += pdf0[bias_sample + i];
        const float f2 = f_square[bias_sample + i];
        path_loss[i] = f2 / pdf;
        path_d[bias_sample + i] = -f2 / pdf / pdf;   // inver_gradient (:832-840)
    }
}
// dE[cell] = sum over the batch's nodes of that cell of peak * d(path), summed in node order.  `order` lists the node indices
// sorted by (batch, cell) (stable, built once per training run); [lo, hi) is this batch's slice.  The reference uses thrust
// sort_by_key + reduce_by_key here (device_thrust.cu:1045-1088), whose summation order is unspecified; ours is fixed, so a
// training run is bit-reproducible.
__global__ void k_train_dE_ordered(const int* __restrict__ order, const uint32_t* __restrict__ sorted_keys, int lo, int hi, const float* __restrict__ peak,
                                   const int* __restrict__ node_path, const float* __restrict__ path_d, const int* __restrict__ label_E, float* __restrict__ dE) {
    const int j = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= hi) return;
    const uint32_t key = sorted_keys[j];
    if (j > lo && sorted_keys[j - 1] == key) return;
    float sum = 0.f;
    int cell = 0;
    for (int k = j; k < hi && sorted_keys[k] == key; k++) {
        const int node = order[k];
        cell = label_E[node];
        sum += peak[node] * path_d[node_path[node]];
    }
    dE[cell] = sum;
}
__global__ void k_node_keys(const int* __restrict__ P2N, int N, int M, int batch, int cells_bits, const int* __restrict__ label_E,
                            uint32_t* __restrict__ keys, int* __restrict__ idx, int* __restrict__ node_path) {
    // one thread per path: its nodes get key (batch index << cells_bits) | cell
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int b = P2N[i], e = (i + 1 < N) ? P2N[i + 1] : M;
    const uint32_t hi = (uint32_t)(i / batch) << cells_bits;
    for (int k = b; k < e; k++) {
        keys[k] = hi | (uint32_t)label_E[k];
        idx[k] = k;
        node_path[k] = i;
    }
}
// fixed-order sum of n values into *out (one block): strided serial partials, then a shared-memory tree
__global__ void k_sum_fixed(const float* __restrict__ v, int n, float* __restrict__ out) {
    __shared__ float s[256];
    float p = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) p += v[i];
    s[threadIdx.x] = p;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s[0];
}
// one block per row: dE_sum, d(loss)/d(theta) (gradient_E2theta, :1090-1147) and the Adam step (:1437-1470)
__global__ void k_train_step(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ E, const float* __restrict__ Esum,
                             const float* __restrict__ dE, int K, int t, float lr, float beta1, float beta2, float eps) {
    __shared__ float s_red[32];
    const int row = blockIdx.x;
    const float den = Esum[row];
    float part = 0.f;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        const float res = E[(size_t)row * K + j];
        const float value = res * den;
        part += (-value / den / den) * dE[(size_t)row * K + j];   // inver_gradient_res * dE
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float x = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (threadIdx.x == 0) s_red[0] = x;
    }
    __syncthreads();
    const float dEsum = s_red[0];
    const float bc1 = 1 - cm_powf(beta1, (float)t), bc2 = 1 - cm_powf(beta2, (float)t);
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        const size_t i = (size_t)row * K + j;
        const float th = theta[i];
        const float sig = sigmoidf_ref(th);
        const float a = sig * (1 - sig) * dEsum;          // sigmoid_gradient_theta * dE_sum
        const float sg = E[i] * den;                        // theta_gradient
        const float b0 = sg * (1 - sg) / den * dE[i];
        const float g = a + b0;
        const float mi = beta1 * m[i] + (1 - beta1) * g;
        const float vi = beta2 * v[i] + (1 - beta2) * (g * g);
        m[i] = mi;
        v[i] = vi;
        const float step = (mi / bc1) / (sqrtf(vi / bc2) + eps);
        if (!isnan(step)) theta[i] = th - lr * step;
    }
}
__global__ void k_theta_init(const float* __restrict__ G, size_t n, float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    theta[i] = (float)(-log(1.0 / (double)G[i] - 1));   // inver_sigmoid (:714-721)
    m[i] = 0.f;
    v[i] = 0.f;
}
// matrix_parameter::toE (:1583-1599): sigmoid + row normalise, no conservative mixing
__global__ void k_train_toE(const float* __restrict__ theta, int K, float* __restrict__ G) {
    __shared__ float s_red[32];
    const int row = blockIdx.x;
    float part = 0.f;
    for (int j = threadIdx.x; j < K; j += blockDim.x) {
        const float s = sigmoidf_ref(theta[(size_t)row * K + j]);
        G[(size_t)row * K + j] = s;
        part += s;
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float x = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (threadIdx.x == 0) s_red[0] = x;
    }
    __syncthreads();
    const float sum = s_red[0];
    for (int j = threadIdx.x; j < K; j += blockDim.x) G[(size_t)row * K + j] /= sum;
}

float* train_optimal_E(Context& c, int batch_size, int epochs, float lr, float* loss_out_host, int loss_cap, int* n_loss) {
    NvtxRange range("spc: train_optimal_E");
    TrainBuffers& t = c.train;
    SPC_REQUIRE(t.N > 0 && t.gamma.p, SPC_ERR_INVALID, "train_optimal_E needs build_optimal_E_train_data and preprocess_getGamma first");
    SPC_REQUIRE(batch_size > 0 && t.N >= batch_size, SPC_ERR_INVALID, "train_optimal_E: batch %d > %d training paths", batch_size, t.N);
    const int K = c.K;
    const size_t n = (size_t)K * K;
    cudaStream_t st = c.stream;
    t.theta.alloc(n); t.adam_m.alloc(n); t.adam_v.alloc(n); t.E.alloc(n); t.dE.alloc(n); t.Esum.alloc(K);
    const int num_batches = t.N / batch_size;
    t.loss.alloc((size_t)num_batches * epochs + 1);
    SPC_CUDA(cudaMemsetAsync(t.loss.p, 0, ((size_t)num_batches * epochs + 1) * sizeof(float), st));
    k_theta_init<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(t.gamma.p, n, t.theta.p, t.adam_m.p, t.adam_v.p);
    c.launches++;
    const float c_keep = (float)(1 - 0.2), c_uniform = (float)(0.2 / (double)(float)K);   // CONSERVATIVE_RATE (optixPathTracer.h:36)
    // nodes sorted by (batch, cell), stable: one sort for the whole run
    const int cells_bits = bits_for(n);
    const int batch_bits = bits_for((size_t)num_batches + 1);
    SPC_REQUIRE(cells_bits + batch_bits <= 32, SPC_ERR_INVALID, "train_optimal_E: K^2 x batches does not fit the 32-bit sort key");
    t.sort_keys.alloc(t.M); t.sort_keys2.alloc(t.M); t.sort_idx.alloc(t.M); t.sort_idx2.alloc(t.M); t.node_path.alloc(t.M);
    t.path_d.alloc(t.N); t.path_loss.alloc(batch_size);
    k_node_keys<<<(t.N + 255) / 256, 256, 0, st>>>(t.P2N.p, t.N, t.M, batch_size, cells_bits, t.label_E.p, t.sort_keys.p, t.sort_idx.p, t.node_path.p);
    {
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, t.sort_keys.p, t.sort_keys2.p, t.sort_idx.p, t.sort_idx2.p, t.M, 0, cells_bits + batch_bits, st);
        t.sort_tmp.alloc(tmp_bytes);
        SPC_CUDA(cub::DeviceRadixSort::SortPairs(t.sort_tmp.p, tmp_bytes, t.sort_keys.p, t.sort_keys2.p, t.sort_idx.p, t.sort_idx2.p, t.M, 0, cells_bits + batch_bits, st));
    }
    c.launches += 2;
    int step = 0;
    for (int ep = 0; ep < epochs; ep++)
        for (int b = 0; b < num_batches; b++) {
            const int bias_sample = b * batch_size;
            step++;
            k_train_E<<<K, 256, 0, st>>>(t.theta.p, K, c_keep, c_uniform, t.E.p, t.Esum.p, t.dE.p);
            k_train_paths<<<(batch_size + 127) / 128, 128, 0, st>>>(t.E.p, t.f_square.p, t.pdf0.p, t.peak.p, t.label_E.p, t.P2N.p, bias_sample, batch_size,
                                                                   t.N, t.M, t.path_d.p, t.path_loss.p);
            // the batch's nodes are the contiguous range [P2N[bias], P2N[bias + batch]) and keep that range after the (batch, cell) sort
            const int lo = t.h_P2N[bias_sample];
            const int hi = bias_sample + batch_size < t.N ? t.h_P2N[bias_sample + batch_size] : t.M;
            if (hi > lo)
                k_train_dE_ordered<<<(hi - lo + 255) / 256, 256, 0, st>>>(t.sort_idx2.p, t.sort_keys2.p, lo, hi, t.peak.p, t.node_path.p, t.path_d.p, t.label_E.p, t.dE.p);
            // sharded training set: this rank's batch is 1/world of the global batch -- sum the K x K gradient over the ranks
            // (NCCL over NVLink, 4 MB per step at K = 1000) so that every rank takes the same Adam step
            comm_allreduce_sum(c, t.dE.p, n);
            k_sum_fixed<<<1, 256, 0, st>>>(t.path_loss.p, batch_size, t.loss.p + (step - 1));
            k_train_step<<<K, 256, 0, st>>>(t.theta.p, t.adam_m.p, t.adam_v.p, t.E.p, t.Esum.p, t.dE.p, K, step, lr, 0.9f, 0.999f, 1e-8f);
            c.launches += 5;
        }
    k_train_toE<<<K, 256, 0, st>>>(t.theta.p, K, t.gamma.p);
    c.launches++;
    SPC_CUDA(cudaGetLastError());
    comm_allreduce_sum(c, t.loss.p, (size_t)step);
    if (n_loss) *n_loss = step;
    if (loss_out_host && loss_cap > 0) {
        const int m = std::min(loss_cap, step);
        std::vector<float> h(m);
        SPC_CUDA(cudaMemcpyAsync(h.data(), t.loss.p, m * sizeof(float), cudaMemcpyDeviceToHost, st));
        SPC_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < m; i++) loss_out_host[i] = h[i] / ((float)batch_size * (float)c.comm_world);
    }
    return t.gamma.p;
}

// ---------------------------------------------------------------------------------------------
// Gamma2CMFGamma: 0.8 * Gamma + 0.2 / K, sequential row prefix, last entry forced to 1
// ---------------------------------------------------------------------------------------------
__global__ void k_gamma_to_cmf(const float* __restrict__ G, int K, float* __restrict__ cmf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K) return;
    const float t = 0.2f;   // CONSERVATIVE_RATE as a float (device_thrust.cu:3415)
    float run = 0.f;
    for (int j = 0; j < K; j++) {
        const float p = (float)((double)(G[(size_t)i * K + j] * (1 - t)) + (1.0 / K) * (double)t);
        run = j == 0 ? p : p + run;
        cmf[(size_t)i * K + j] = run;
    }
    cmf[(size_t)(i + 1) * K - 1] = 1;
}
// Guide tables of the CDF rows (shade.cuh, "guide tables").  The CDF travels to the eye pass as a bare pointer inside MyParams, possibly
// through another context on the same device (frame lanes share the trained state by pointer), so the tables are found through a
// process-wide registry keyed by that pointer; a CDF this library did not build has no entry and is sampled with the reference's bisect.
__global__ void k_gamma_guide(const float* __restrict__ cmf, int K, int* __restrict__ guide) {
    const int row = blockIdx.x;
    if (row < K) guide_build_table(cmf + (size_t)row * K, K, guide + (size_t)row * (K + 1));
}
namespace {
struct GammaGuide {
    const int* guide;
    int        K;
    const void* owner;
};
std::mutex g_guide_mutex;
std::map<const float*, GammaGuide> g_gamma_guides;
struct CTreeEntry {
    const float4* ctree;
    const void*   owner;
};
std::map<const spc_tree_node*, CTreeEntry> g_ctrees;
}  // namespace
const float4* ctree_lookup(const spc_tree_node* tree_dev) {
    if (!tree_dev) return nullptr;
    std::lock_guard<std::mutex> lock(g_guide_mutex);
    auto it = g_ctrees.find(tree_dev);
    return it != g_ctrees.end() ? it->second.ctree : nullptr;
}
void ctree_register(const spc_tree_node* tree_dev, const float4* ctree_dev, const void* owner) {
    std::lock_guard<std::mutex> lock(g_guide_mutex);
    if (ctree_dev) g_ctrees[tree_dev] = CTreeEntry{ctree_dev, owner};
    else g_ctrees.erase(tree_dev);
}
const int* gamma_guide_lookup(const float* cmf_gamma, int K) {
    std::lock_guard<std::mutex> lock(g_guide_mutex);
    auto it = g_gamma_guides.find(cmf_gamma);
    return (it != g_gamma_guides.end() && it->second.K == K) ? it->second.guide : nullptr;
}
void gamma_guide_forget(const void* owner, bool trees_too) {
    std::lock_guard<std::mutex> lock(g_guide_mutex);
    for (auto it = g_gamma_guides.begin(); it != g_gamma_guides.end();)
        it = (it->second.owner == owner) ? g_gamma_guides.erase(it) : std::next(it);
    if (!trees_too) return;
    for (auto it = g_ctrees.begin(); it != g_ctrees.end();)
        it = (it->second.owner == owner) ? g_ctrees.erase(it) : std::next(it);
}
float* train_gamma_to_cmf(Context& c, const float* gamma_dev) {
    TrainBuffers& t = c.train;
    const int K = c.K;
    gamma_guide_forget(&c, false);
    t.cmf.alloc((size_t)K * K);
    t.cmf_guide.alloc((size_t)K * (K + 1));
    k_gamma_to_cmf<<<(K + 63) / 64, 64, 0, c.stream>>>(gamma_dev, K, t.cmf.p);
    k_gamma_guide<<<K, 128, 0, c.stream>>>(t.cmf.p, K, t.cmf_guide.p);
    c.launches += 2;
    SPC_CUDA(cudaGetLastError());
    {
        // contexts on other streams may use the tables (frame lanes): make them complete before the pointer is handed out
        SPC_CUDA(cudaStreamSynchronize(c.stream));
        std::lock_guard<std::mutex> lock(g_guide_mutex);
        g_gamma_guides[t.cmf.p] = GammaGuide{t.cmf_guide.p, K, &c};
    }
    return t.cmf.p;
}

}  // namespace spc
