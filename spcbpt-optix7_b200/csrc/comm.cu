// comm.cu -- the multi-GPU exchanges of the render path, NCCL over NVLink / NVSwitch, behind the C ABI.
// One process per GPU; every rank holds a replica of scene + BVH and renders its own subframes.  The path has exactly two exchange
// steps, and NCCL appears nowhere else:
//   * subspace training, once: the training paths and the Q light-trace launches are SHARDED across ranks, and the statistics
//     derived from them are all-reduced where the single-GPU code sums over the whole set -- the 10x10-pixel reweighting grid
//     (sample_reweight), Q (spc_allreduce_training_stats), the Gamma histogram before its row normalisation, and the K x K gradient
//     of every Adam step right after k_train_dE_ordered (train.cu); trees are built on rank 0's host and broadcast;
//   * read-out: the accumulation buffers are reduced to the root (spc_reduce_accum).
// The reference is single-GPU (no counterpart); sutil/WorkDistribution.h:34-91 is its unused multi-GPU tile partition.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include "common.cuh"

namespace spc {

// NCCL is bound at the first spc_comm_* call, not at link time: a process may already hold a libnccl.so.2 (PyTorch ships its own and
// resolves it by the same SONAME -- linking ours eagerly would make whichever library loads first win for both), and single-GPU
// users need none at all.  RTLD_NOLOAD first: reuse the copy the process has; else the system library.
namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
};
NcclApi g_nccl = {};
bool g_nccl_ready = false;
std::mutex g_nccl_mutex;
}  // namespace

static const NcclApi& nccl_api() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl_ready) return g_nccl;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        set_error("multi-GPU needs NCCL: cannot load libnccl.so.2 (%s)", dlerror());
        throw CudaFailure{SPC_ERR_INVALID};
    }
    auto sym = [&](const char* name) {
        void* p = dlsym(h, name);
        if (!p) {
            set_error("libnccl.so.2 lacks %s", name);
            throw CudaFailure{SPC_ERR_INVALID};
        }
        return p;
    };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
    g_nccl.Reduce = reinterpret_cast<decltype(g_nccl.Reduce)>(sym("ncclReduce"));
    g_nccl.Broadcast = reinterpret_cast<decltype(g_nccl.Broadcast)>(sym("ncclBroadcast"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    g_nccl_ready = true;
    return g_nccl;
}
// the calls below read like plain NCCL
#define ncclGetUniqueId spc::nccl_api().GetUniqueId
#define ncclCommInitRank spc::nccl_api().CommInitRank
#define ncclCommDestroy spc::nccl_api().CommDestroy
#define ncclAllReduce spc::nccl_api().AllReduce
#define ncclReduce spc::nccl_api().Reduce
#define ncclBroadcast spc::nccl_api().Broadcast
#define ncclGetErrorString spc::nccl_api().GetErrorString

#define SPC_NCCL(call)                                                                                   \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) {                                                                        \
            spc::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, ncclGetErrorString(r__)); \
            throw spc::CudaFailure{SPC_ERR_CUDA};                                                        \
        }                                                                                                \
    } while (0)

static ncclComm_t nccl(Context& c) { return static_cast<ncclComm_t>(c.comm); }

void comm_allreduce_sum(Context& c, float* dev, size_t n) {
    if (c.comm_world <= 1 || n == 0) return;
    SPC_NCCL(ncclAllReduce(dev, dev, n, ncclFloat, ncclSum, nccl(c), c.stream));
}
void comm_bcast(Context& c, void* dev, size_t bytes, int root) {
    if (c.comm_world <= 1 || bytes == 0) return;
    SPC_NCCL(ncclBroadcast(dev, dev, bytes, ncclChar, root, nccl(c), c.stream));
}

__global__ void k_scale4(spc_float4* __restrict__ a, int n, float w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    spc_float4 v = a[i];
    a[i] = spc_float4{v.x * w, v.y * w, v.z * w, v.w * w};
}

}  // namespace spc

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

int spc_comm_unique_id(void* id_out) {
    static_assert(sizeof(ncclUniqueId) == SPC_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id_out) {
        spc::set_error("spc_comm_unique_id: null output");
        return SPC_ERR_INVALID;
    }
    ncclUniqueId id;
    const ncclResult_t r = ncclGetUniqueId(&id);
    if (r != ncclSuccess) {
        spc::set_error("ncclGetUniqueId: %s", ncclGetErrorString(r));
        return SPC_ERR_CUDA;
    }
    memcpy(id_out, &id, sizeof(id));
    return SPC_OK;
}

int spc_comm_init(spc_context* ctx, int rank, int world, const void* id) {
    SPC_API_BEGIN
    SPC_REQUIRE(world >= 1 && rank >= 0 && rank < world && (world == 1 || id), SPC_ERR_INVALID, "spc_comm_init: rank %d of %d", rank, world);
    SPC_REQUIRE(!c.comm, SPC_ERR_INVALID, "spc_comm_init: the context already has a communicator");
    if (world > 1) {
        ncclUniqueId uid;
        memcpy(&uid, id, sizeof(uid));
        ncclComm_t comm = nullptr;
        SPC_NCCL(ncclCommInitRank(&comm, world, uid, rank));
        c.comm = comm;
        // NCCL sets its channels up lazily, on the first collective (~1 s with 8 ranks): pay for it here, not inside the training
        c.comm_scratch.alloc(4);
        SPC_CUDA(cudaMemsetAsync(c.comm_scratch.p, 0, 4, c.stream));
        SPC_NCCL(ncclAllReduce(c.comm_scratch.p, c.comm_scratch.p, 1, ncclInt32, ncclSum, comm, c.stream));
        SPC_CUDA(cudaStreamSynchronize(c.stream));
    }
    c.comm_rank = rank;
    c.comm_world = world;
    SPC_API_END
}

int spc_comm_destroy(spc_context* ctx) {
    SPC_API_BEGIN
    if (c.comm) {
        SPC_CUDA(cudaStreamSynchronize(c.stream));
        ncclCommDestroy(static_cast<ncclComm_t>(c.comm));
        c.comm = nullptr;
    }
    c.comm_rank = 0;
    c.comm_world = 1;
    SPC_API_END
}

int spc_comm_info(spc_context* ctx, int* rank, int* world) {
    SPC_API_BEGIN
    if (rank) *rank = c.comm_rank;
    if (world) *world = c.comm_world;
    SPC_API_END
}

// small host-side values (path counts, step counts, timings): staged through a device scratch buffer
int spc_comm_allreduce_host(spc_context* ctx, void* host_buf, int count, int dtype, int op) {
    SPC_API_BEGIN
    SPC_REQUIRE(host_buf && count > 0 && count <= 4096 && dtype >= 0 && dtype <= 2 && op >= 0 && op <= 2, SPC_ERR_INVALID, "spc_comm_allreduce_host: bad arguments");
    if (c.comm_world > 1) {
        const size_t es = dtype == SPC_COMM_F64 ? 8 : 4;
        c.comm_scratch.alloc(count * es);
        SPC_CUDA(cudaMemcpyAsync(c.comm_scratch.p, host_buf, count * es, cudaMemcpyHostToDevice, c.stream));
        const ncclDataType_t dt = dtype == SPC_COMM_I32 ? ncclInt32 : (dtype == SPC_COMM_F32 ? ncclFloat32 : ncclFloat64);
        const ncclRedOp_t ro = op == SPC_COMM_SUM ? ncclSum : (op == SPC_COMM_MIN ? ncclMin : ncclMax);
        SPC_NCCL(ncclAllReduce(c.comm_scratch.p, c.comm_scratch.p, count, dt, ro, static_cast<ncclComm_t>(c.comm), c.stream));
        SPC_CUDA(cudaMemcpyAsync(host_buf, c.comm_scratch.p, count * es, cudaMemcpyDeviceToHost, c.stream));
        SPC_CUDA(cudaStreamSynchronize(c.stream));
    }
    SPC_API_END
}

int spc_comm_bcast_host(spc_context* ctx, void* host_buf, size_t bytes, int root) {
    SPC_API_BEGIN
    SPC_REQUIRE((host_buf || !bytes) && root >= 0 && root < c.comm_world, SPC_ERR_INVALID, "spc_comm_bcast_host: bad arguments");
    if (c.comm_world > 1 && bytes) {
        c.comm_scratch.alloc(bytes);
        if (c.comm_rank == root) SPC_CUDA(cudaMemcpyAsync(c.comm_scratch.p, host_buf, bytes, cudaMemcpyHostToDevice, c.stream));
        spc::comm_bcast(c, c.comm_scratch.p, bytes, root);
        if (c.comm_rank != root) SPC_CUDA(cudaMemcpyAsync(host_buf, c.comm_scratch.p, bytes, cudaMemcpyDeviceToHost, c.stream));
        SPC_CUDA(cudaStreamSynchronize(c.stream));
    }
    SPC_API_END
}

int spc_comm_barrier(spc_context* ctx) {
    SPC_API_BEGIN
    SPC_CUDA(cudaStreamSynchronize(c.stream));
    if (c.comm_world > 1) {
        c.comm_scratch.alloc(4);
        SPC_NCCL(ncclAllReduce(c.comm_scratch.p, c.comm_scratch.p, 1, ncclInt32, ncclSum, static_cast<ncclComm_t>(c.comm), c.stream));
        SPC_CUDA(cudaStreamSynchronize(c.stream));
    }
    SPC_API_END
}

// Q of every rank = running mean over ITS light-trace launches, weighted by path counts (preprocess_getQ).  The estimate over all
// ranks' launches is the path-count-weighted mean of the per-rank estimates: Q <- sum_r n_r Q_r / sum_r n_r.
int spc_allreduce_training_stats(spc_context* ctx) {
    SPC_API_BEGIN
    spc::train_allreduce_Q(c);
    SPC_API_END
}

// accum <- weight * accum, then summed over the ranks: into `root`'s buffer (root >= 0; the other ranks' buffers are left scaled),
// or into every rank's buffer (root < 0).  With weight = this rank's share of the subframes the result is the running mean of the
// whole sample-partitioned render (each rank keeps the running mean of ITS subframes, raygen.cu:430-437).
int spc_reduce_accum(spc_context* ctx, spc_float4* accum_dev, int n_pixels, float weight, int root) {
    SPC_API_BEGIN
    SPC_REQUIRE(accum_dev && n_pixels > 0 && root < c.comm_world, SPC_ERR_INVALID, "spc_reduce_accum: bad arguments");
    spc::k_scale4<<<(n_pixels + 255) / 256, 256, 0, c.stream>>>(accum_dev, n_pixels, weight);
    c.launches++;
    SPC_CUDA(cudaGetLastError());
    if (c.comm_world > 1) {
        ncclComm_t comm = static_cast<ncclComm_t>(c.comm);
        if (root < 0) SPC_NCCL(ncclAllReduce(accum_dev, accum_dev, (size_t)n_pixels * 4, ncclFloat, ncclSum, comm, c.stream));
        else SPC_NCCL(ncclReduce(accum_dev, accum_dev, (size_t)n_pixels * 4, ncclFloat, ncclSum, root, comm, c.stream));
    }
    SPC_API_END
}

}  // extern "C"
