// lvc.cu -- light-vertex-cache binning on the device.  Replaces MyThrustOp::LVC_Process
// (cuda_thrust/device_thrust.cu:241-332), which copies the validity flags, subspace ids and weights of all
// 800 k LVC slots to the host, buckets and prefix-sums them in a serial host loop and uploads three arrays
// again every frame.  Here the same result is produced by six small kernels without leaving the GPU:
//
//   k_lvc_keys     per slot: key = subspace id (or -1 when invalid), weight = (flux.x+flux.y+flux.z)/pdf with
//                  Inf/NaN -> 0 (device_thrust.cu:200-207)
//   k_bin_hist     per-chunk histogram of the keys in shared memory
//   k_lvc_colscan  per subspace: exclusive scan of its counts over the chunks (stable order = slot order)
//   k_lvc_bias     exclusive scan over subspaces -> Subspace{jump_bias,id,size}; vertex_count
//   k_lvc_scatter  stable counting-sort scatter (warp match-any ranks) -> jump_buffer + sorted weights
//   k_lvc_cmf      per subspace (one warp): running fp32 sum IN SLOT ORDER -- the summation order of the
//                  reference's host loop, so the cmf values are bit-identical -- then division by the total
#include "shade.cuh"

namespace spc {

constexpr int kChunk = 2048;   // slots per warp-chunk

// per slot: bin key + weight of an LVC vertex; counters[1] += #valid depth-0 vertices (path_count)
__global__ void k_lvc_keys(const spc_vertex* __restrict__ lvc, const uint8_t* __restrict__ valid, int n, int K,
                           int* __restrict__ key, float* __restrict__ weight, int* __restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n_begin = 0;
    if (i < n) {
        int k = -1;
        float w = 0.f;
        if (valid[i]) {
            const spc_vertex* v = lvc + i;
            const int sid = v->subspaceId;
            if (sid >= 0 && sid < K) {
                k = sid;
                float res = (v->flux.x + v->flux.y + v->flux.z) / v->pdf;
                res = isinf(res) ? 0 : res;
                w = isnan(res) ? 0 : res;
                if (v->depth == 0) n_begin = 1;
            }
        }
        key[i] = k;
        weight[i] = w;
    }
    for (int o = 16; o > 0; o >>= 1) n_begin += __shfl_xor_sync(0xffffffffu, n_begin, o);
    if ((threadIdx.x & 31) == 0 && n_begin) atomicAdd(counters + 1, n_begin);
}

// generic: per-chunk histogram of bin keys (key < 0: not binned) in shared memory, one warp per chunk
__global__ void k_bin_hist(const int* __restrict__ key, int n, int K, int n_chunks, int* __restrict__ hist) {
    extern __shared__ int s_hist[];   // [warps][K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    int* h = s_hist + (size_t)warp * K;
    for (int chunk = blockIdx.x * wpb + warp; chunk < n_chunks; chunk += gridDim.x * wpb) {
        for (int k = lane; k < K; k += 32) h[k] = 0;
        __syncwarp();
        const int lo = chunk * kChunk, hi = min(n, lo + kChunk);
        for (int i = lo + lane; i < hi; i += 32) {
            const int k = key[i];
            if (k >= 0) atomicAdd(h + k, 1);
        }
        __syncwarp();
        for (int k = lane; k < K; k += 32) hist[(size_t)chunk * K + k] = h[k];
        __syncwarp();
    }
}

// hist[chunk][k] -> exclusive prefix over chunks (in place); totals[k] = size of subspace k.
// One block per 32 subspaces, 32 x 32 threads: a tile of 32 chunks x 32 subspaces is loaded with coalesced rows (a warp = one chunk row),
// 32 threads scan their subspace's column of the tile in shared memory, the tile is written back the same way; the carry per subspace runs
// over the ~400 chunk rows tile by tile.  (One thread per subspace walking all chunks serially took 100 us of dependent global loads.)
constexpr int kColTile = 32;
__global__ void __launch_bounds__(kColTile * 32) k_lvc_colscan(int* __restrict__ hist, int n_chunks, int K, int* __restrict__ totals) {
    __shared__ int s_tile[kColTile][33];
    __shared__ int s_carry[32];
    const int kx = threadIdx.x & 31, cx = threadIdx.x >> 5;
    const int k = blockIdx.x * 32 + kx;
    if (cx == 0) s_carry[kx] = 0;
    for (int c0 = 0; c0 < n_chunks; c0 += kColTile) {
        const int c = c0 + cx;
        const bool in = c < n_chunks && k < K;
        s_tile[cx][kx] = in ? hist[(size_t)c * K + k] : 0;
        __syncthreads();
        if (cx == 0) {
            int run = s_carry[kx];
#pragma unroll 8
            for (int j = 0; j < kColTile; j++) {
                const int v = s_tile[j][kx];
                s_tile[j][kx] = run;
                run += v;
            }
            s_carry[kx] = run;
        }
        __syncthreads();
        if (in) hist[(size_t)c * K + k] = s_tile[cx][kx];
        __syncthreads();
    }
    if (cx == 0 && k < K) totals[k] = s_carry[kx];
}

// one block: exclusive scan of totals -> Subspace records; counters[0] = vertex_count
__global__ void k_lvc_bias(const int* __restrict__ totals, int K, spc_subspace* __restrict__ sub, int* __restrict__ counters) {
    __shared__ int s_part[1024];
    const int t = threadIdx.x, T = blockDim.x;
    const int per = (K + T - 1) / T;
    const int lo = min(K, t * per), hi = min(K, lo + per);
    int s = 0;
    for (int k = lo; k < hi; k++) s += totals[k];
    s_part[t] = s;
    __syncthreads();
    if (t == 0) {
        int run = 0;
        for (int i = 0; i < T; i++) { const int v = s_part[i]; s_part[i] = run; run += v; }
        counters[0] = run;
    }
    __syncthreads();
    int run = s_part[t];
    for (int k = lo; k < hi; k++) {
        sub[k].jump_bias = run;
        sub[k].id = k;
        sub[k].size = totals[k];
        sub[k].Q = 0.f;
        run += totals[k];
    }
}

__global__ void k_lvc_scatter(const int* __restrict__ key, const float* __restrict__ weight, int n, int K, int n_chunks,
                              const int* __restrict__ hist, const spc_subspace* __restrict__ sub, int* __restrict__ jump,
                              float* __restrict__ wsorted) {
    extern __shared__ int s_cur[];   // [warps][K] write cursors
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    int* cur = s_cur + (size_t)warp * K;
    for (int chunk = blockIdx.x * wpb + warp; chunk < n_chunks; chunk += gridDim.x * wpb) {
        for (int k = lane; k < K; k += 32) cur[k] = sub[k].jump_bias + hist[(size_t)chunk * K + k];
        __syncwarp();
        const int lo = chunk * kChunk, hi = min(n, lo + kChunk);
        for (int base = lo; base < hi; base += 32) {
            const int i = base + lane;
            const int k = i < hi ? key[i] : -1;
            const unsigned active = __ballot_sync(0xffffffffu, k >= 0);
            if (k >= 0) {
                const unsigned peers = __match_any_sync(active, k);
                const int rank = __popc(peers & ((1u << lane) - 1u));
                const int pos = cur[k] + rank;
                jump[pos] = i;
                wsorted[pos] = weight[i];
                __syncwarp(active);
                if (rank == 0) cur[k] += __popc(peers);
            }
            __syncwarp();
        }
    }
}

// one warp per subspace: running fp32 sum in slot order, then normalise.  The summation order is the reference's
// (sequential), so the only serial resource is the FADD dependency chain: the warp stages 256 weights at a time in
// shared memory with coalesced loads, lane 0 runs the chain over shared memory (4-cycle adds, loads off the critical
// path), and all lanes write the prefixes back coalesced.
constexpr int kCmfTile = 256;
__global__ void __launch_bounds__(128) k_lvc_cmf(spc_subspace* __restrict__ sub, int K, const float* __restrict__ wsorted, float* __restrict__ cmfs) {
    __shared__ float s_buf[4][kCmfTile];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = blockIdx.x * 4 + wib;
    if (warp >= K) return;
    float* buf = s_buf[wib];
    const int bias = sub[warp].jump_bias, size = sub[warp].size;
    float run = 0.f;
    for (int base = 0; base < size; base += kCmfTile) {
        const int m = min(kCmfTile, size - base);
        for (int j = lane; j < m; j += 32) buf[j] = wsorted[bias + base + j];
        __syncwarp();
        if (lane == 0) {
            // the reference's first element is the weight itself, later ones `w + previous` (device_thrust.cu:281-286)
            int j = 0;
            if (base == 0) { run = buf[0]; j = 1; }
#pragma unroll 8
            for (; j < m; j++) {
                run = buf[j] + run;
                buf[j] = run;
            }
        }
        __syncwarp();
        for (int j = lane; j < m; j += 32) cmfs[bias + base + j] = buf[j];
        __syncwarp();
    }
    // Q_subspace_vertex[s] += w accumulates the same sequence starting from 0: 0 + w0 = w0, so it equals the last running sum
    const float total = __shfl_sync(0xffffffffu, size > 0 ? run : 0.f, 0);
    if (lane == 0) sub[warp].sum_pmf = total;
    for (int i = lane; i < size; i += 32) cmfs[bias + i] = cmfs[bias + i] / total;
}

// guide tables of the per-subspace cmfs (shade.cuh, "guide tables"): one block per subspace
__global__ void k_lvc_guide(const spc_subspace* __restrict__ sub, int K, const float* __restrict__ cmfs, int* __restrict__ guide) {
    const int b = blockIdx.x;
    if (b >= K) return;
    const int n = sub[b].size, bias = sub[b].jump_bias;
    if (n > 0) guide_build_table(cmfs + bias, n, guide + bias + b);
}

// Generic ordered binning: given per-element bin keys (-1 = skip) and weights, produce the stable bucket order
// (jump), Subspace{jump_bias,id,size,sum_pmf} per bin with sum_pmf = the fp32 sum of the bin's weights IN ELEMENT
// ORDER, and (optionally normalised) running sums.  Used by LVC_Process, preprocess_getQ and sample_reweight,
// whose reference implementations are serial host loops with exactly this summation order.
void bin_ordered(Context& c, LvcBuffers& b, int n, int K, int* counters /* [0] <- number of binned elements */) {
    const int n_chunks = (n + kChunk - 1) / kChunk;
    b.guide_valid = false;   // the cmfs change: lvc_process rebuilds the guide tables
    b.subspace.alloc(K); b.cmfs.alloc(n); b.jump.alloc(n); b.wsorted.alloc(n);
    b.hist.alloc((size_t)n_chunks * K);
    cudaStream_t st = c.stream;
    // shared memory: K ints per warp; as many warps per block as fit in 160 KB (K = 1000 -> 8 warps, 32 KB)
    const int wpb = (int)std::min<size_t>(8, (160 * 1024) / ((size_t)K * 4));
    SPC_REQUIRE(wpb >= 1, SPC_ERR_CAPACITY, "ordered binning: %d bins do not fit the shared-memory histogram", K);
    const size_t smem = (size_t)wpb * K * 4;
    if (!c.lvc_attr_set) {   // function attributes are per device: once per context (a context is driven by one host thread at a time)
        SPC_CUDA(cudaFuncSetAttribute(k_bin_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SPC_CUDA(cudaFuncSetAttribute(k_lvc_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        c.lvc_attr_set = true;
    }
    const int grid = std::max(1, std::min((n_chunks + wpb - 1) / wpb, c.sm_count * 4));
    k_bin_hist<<<grid, wpb * 32, smem, st>>>(b.key.p, n, K, n_chunks, b.hist.p);
    k_lvc_colscan<<<(K + 31) / 32, kColTile * 32, 0, st>>>(b.hist.p, n_chunks, K, b.totals.p);
    k_lvc_bias<<<1, 1024, 0, st>>>(b.totals.p, K, b.subspace.p, counters);
    k_lvc_scatter<<<grid, wpb * 32, smem, st>>>(b.key.p, b.weight.p, n, K, n_chunks, b.hist.p, b.subspace.p, b.jump.p, b.wsorted.p);
    k_lvc_cmf<<<(K + 3) / 4, 128, 0, st>>>(b.subspace.p, K, b.wsorted.p, b.cmfs.p);
    c.launches += 5;
    SPC_CUDA(cudaGetLastError());
}

// keys + weights of an LVC, then ordered binning; counters (device, 2 ints after totals) = {vertex_count, path_count}
int* lvc_bin(Context& c, LvcBuffers& b, const spc_vertex* lvc, const uint8_t* valid, int n) {
    const int K = c.K;
    b.weight.alloc(n); b.key.alloc(n); b.totals.alloc(K + 8);
    b.n = n;
    int* counters = b.totals.p + K;
    SPC_CUDA(cudaMemsetAsync(counters, 0, 8 * sizeof(int), c.stream));
    k_lvc_keys<<<(n + 255) / 256, 256, 0, c.stream>>>(lvc, valid, n, K, b.key.p, b.weight.p, counters);
    c.launches++;
    bin_ordered(c, b, n, K, counters);
    return counters;
}

void lvc_process(Context& c, const spc_vertex* lvc, const uint8_t* valid, int n, spc_subspace_sampler* out) {
    NvtxRange range("spc: LVC_Process");
    SPC_REQUIRE(lvc && valid && n > 0 && out, SPC_ERR_INVALID, "spc_lvc_process: bad arguments");
    LvcBuffers& b = c.lvc;
    if (!c.h_pinned) SPC_CUDA(cudaMallocHost((void**)&c.h_pinned, 64 * sizeof(int)));
    int* counters = lvc_bin(c, b, lvc, valid, n);
    b.guide.alloc((size_t)n + c.K);
    k_lvc_guide<<<c.K, 128, 0, c.stream>>>(b.subspace.p, c.K, b.cmfs.p, b.guide.p);
    c.launches++;
    b.guide_valid = true;
    SPC_CUDA(cudaMemcpyAsync(c.h_pinned, counters, 2 * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    out->LVC = lvc;
    out->subspace = b.subspace.p;
    out->cmfs = b.cmfs.p;
    out->jump_buffer = b.jump.p;
    out->vertex_count = c.h_pinned[0];
    out->path_count = c.h_pinned[1];
}

}  // namespace spc
