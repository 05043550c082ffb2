// spcbpt_main.cpp -- the host driver: the reference application's schedule on top of the C ABI of libspcbpt_b200.so.
//
// Stands in for src/OptiXPathTracer/optixPathTracer.cpp without the GLFW/GL/ImGui front end: scene loading
// (LoadScene + Scene_shift + LightSource_shift, :735-741), buffer set-up (initLaunchParams :260-310, lt_params_setup
// :462-476, preTracer_params_setup :479-488), the subspace-training schedule (preprocessing :552-608) and the
// per-frame loop (updateState :371-380, launchLVCTrace :515-522, launchSubframe :609-635, main :790-822), with the
// reference's timing read-out (sutil::displayStats) replaced by a summary line.  Every GPU step is one call into the
// library; this file only sequences them (the Python twin is spcbpt-optix7_b200/renderer.py, used by the tests).
//
// Display is replaced by files: <out>.ppm / <out>.png (tone-mapped frame buffer) and <out>.pfm (accumulation buffer).
#include <cuda_runtime_api.h>
#include <signal.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <fstream>
#include <exception>
#include <memory>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "host_scene.hpp"
#include "spcbpt_b200.h"
#include "train_state.hpp"

using spchost::HostScene;

#define SPC_CHECK(call)                                                                                   \
    do {                                                                                                  \
        if ((call) != SPC_OK) throw std::runtime_error(std::string(#call) + ": " + spc_last_error());     \
    } while (0)
#define CUDA_CHECK(call)                                                                                  \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

namespace {

struct Options {
    std::string scene, cache, data_root, out = "spcbpt_out", save_cache, alg = "SPCBPT_eye", save_state, load_state;
    int  width = 1920, height = 1000;   // optixPathTracer.cpp:700-701
    int  frames = 16, device = 0, lanes = 1;
    int  K = 1000, K_light = 0, connections = 3, max_depth = 0;
    int  train_samples = 2000000, q_samples = 2000000, tree_samples = 100000, batch = 20000, epochs = 1;
    float lr = 0.01f;
    int  lt_cores = 1000, lt_padding = 800, lt_per_core = 100;   // lt_params_setup
    int  pre_cores = 10000, pre_padding = 10;                    // preTracer_params_setup
    unsigned seed_offset = 0, seed_stride = 1;
    bool pipeline = true, render = true, quiet = false, write_images = true;
    bool host_trees = false;   // classification trees by the host builder (the reference's way) instead of the device-side build
    // multi-GPU: one process per GPU.  `--ranks N` forks N ranks on devices 0..N-1 of this node; or start the ranks yourself
    // with --rank r --world N --id-file <path> (rank 0 writes the NCCL unique id there, the others wait for it)
    int  ranks = 1, rank = 0, world = 1;
    bool tiles = false;   // multi-GPU by image tiles (spc_set_tile_partition) instead of by samples: frame latency scales, not throughput
    std::string id_file;
    std::vector<std::pair<std::string, long long>> options;   // --option name=value -> spc_set_option on every context
};

void usage(const char* argv0) {
    fprintf(stderr,
            "Usage  : %s --scene <file.scene> | --cache <file.spcscene> [options]\n"
            "         --dim=<width>x<height>      image dimensions; defaults to 1920x1000\n"
            "         --no-gl-interop             accepted for command-line compatibility with the reference (output goes to files)\n"
            "         --data-root <dir>           directory the .scene's file names are relative to\n"
            "         --frames <n>                subframes to accumulate per rank (default 16)\n"
            "         --host-trees                build the classification trees on the host (the reference's way; default: on the GPU)\n"
            "         --option name=value         context switch (spc_set_option): light_trace_mode=1, tail_threshold=-1, ...\n"
            "         --ranks <n>                 multi-GPU: fork n ranks on devices 0..n-1 (NCCL inside the library: sharded training,\n"
            "                                     sample-partitioned frames, accumulation buffer reduced to rank 0)\n"
            "         --rank r --world n --id-file f   the same with externally started ranks (rank 0 writes the NCCL id to f)\n"
            "         --tiles                     with several ranks: partition every frame by 8x4 pixel tiles (sutil/WorkDistribution.h) instead of\n"
            "                                     partitioning the samples: all ranks render the same subframes, each its own tiles\n"
            "         --alg pt|SPCBPT_eye         integrator (the reference toggles these with Space)\n"
            "         --K <n> --K-light <n> --connections <n> --max-depth <n>\n"
            "         --train-samples <n> --q-samples <n> --tree-samples <n> --batch <n> --epochs <n> --lr <f>\n"
            "         --lt-cores <n> --lt-padding <n> --lt-per-core <n> --pretrace-cores <n> --pretrace-padding <n>\n"
            "         --lanes <n>                 frame lanes: n contexts on their own streams and host threads render alternate subframes\n"
            "                                     concurrently (same samples as the sequential loop, running means merged at read-out)\n"
            "         --device <i> --seed-offset <u> --seed-stride <u> --no-pipeline --no-images --quiet\n"
            "         --out <prefix>              writes <prefix>.ppm and <prefix>.pfm (default spcbpt_out)\n"
            "         --save-cache <file>         write the parsed scene as a .spcscene cache\n"
            "         --save-state <prefix>       write trees, Q and Gamma as <prefix>tree_eye.txt, tree_light.txt, Q.txt, E.txt\n"
            "         --load-state <prefix>       read them back instead of training (the reference's tree_load / load_Q_file / load_Gamma_file)\n"
            "         --no-render                 stop after loading (with --save-cache: scene conversion only, no GPU)\n",
            argv0);
}

// sutil::parseDimensions (sutil/sutil.cpp:768-793)
bool parse_dimensions(const char* arg, int& w, int& h) {
    const char* x = strchr(arg, 'x');
    if (!x || x == arg || !x[1]) return false;
    w = atoi(std::string(arg, x).c_str());
    h = atoi(x + 1);
    return w > 0 && h > 0;
}

bool parse_args(int argc, char** argv, Options& o) {
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto need = [&](const char* name) -> const char* {
            if (i + 1 >= argc) throw std::runtime_error(std::string("missing value for ") + name);
            return argv[++i];
        };
        if (a == "--help" || a == "-h") return false;
        else if (a == "--no-gl-interop") {}   // the reference's third option (optixPathTracer.cpp:702): accepted, there is no GL interop to disable
        else if (a.rfind("--dim=", 0) == 0) { if (!parse_dimensions(a.c_str() + 6, o.width, o.height)) throw std::runtime_error("Failed to parse width, height from string '" + a.substr(6) + "'"); }
        else if (a == "--scene") o.scene = need("--scene");
        else if (a == "--cache") o.cache = need("--cache");
        else if (a == "--data-root") o.data_root = need("--data-root");
        else if (a == "--out") o.out = need("--out");
        else if (a == "--save-cache") o.save_cache = need("--save-cache");
        else if (a == "--alg") o.alg = need("--alg");
        else if (a == "--save-state") o.save_state = need("--save-state");
        else if (a == "--load-state") o.load_state = need("--load-state");
        else if (a == "--frames") o.frames = atoi(need("--frames"));
        else if (a == "--device") o.device = atoi(need("--device"));
        else if (a == "--lanes") o.lanes = atoi(need("--lanes"));
        else if (a == "--K") o.K = atoi(need("--K"));
        else if (a == "--K-light") o.K_light = atoi(need("--K-light"));
        else if (a == "--connections") o.connections = atoi(need("--connections"));
        else if (a == "--max-depth") o.max_depth = atoi(need("--max-depth"));
        else if (a == "--train-samples") o.train_samples = atoi(need("--train-samples"));
        else if (a == "--q-samples") o.q_samples = atoi(need("--q-samples"));
        else if (a == "--tree-samples") o.tree_samples = atoi(need("--tree-samples"));
        else if (a == "--batch") o.batch = atoi(need("--batch"));
        else if (a == "--epochs") o.epochs = atoi(need("--epochs"));
        else if (a == "--lr") o.lr = (float)atof(need("--lr"));
        else if (a == "--lt-cores") o.lt_cores = atoi(need("--lt-cores"));
        else if (a == "--lt-padding") o.lt_padding = atoi(need("--lt-padding"));
        else if (a == "--lt-per-core") o.lt_per_core = atoi(need("--lt-per-core"));
        else if (a == "--pretrace-cores") o.pre_cores = atoi(need("--pretrace-cores"));
        else if (a == "--pretrace-padding") o.pre_padding = atoi(need("--pretrace-padding"));
        else if (a == "--seed-offset") o.seed_offset = (unsigned)strtoul(need("--seed-offset"), nullptr, 10);
        else if (a == "--seed-stride") o.seed_stride = (unsigned)strtoul(need("--seed-stride"), nullptr, 10);
        else if (a == "--option") {
            const std::string kv = need("--option");
            const size_t eq = kv.find('=');
            if (eq == std::string::npos) throw std::runtime_error("--option expects name=value");
            o.options.emplace_back(kv.substr(0, eq), atoll(kv.c_str() + eq + 1));
        }
        else if (a == "--ranks") o.ranks = atoi(need("--ranks"));
        else if (a == "--rank") o.rank = atoi(need("--rank"));
        else if (a == "--world") o.world = atoi(need("--world"));
        else if (a == "--id-file") o.id_file = need("--id-file");
        else if (a == "--tiles") o.tiles = true;
        else if (a == "--host-trees") o.host_trees = true;
        else if (a == "--no-pipeline") o.pipeline = false;
        else if (a == "--no-render") o.render = false;
        else if (a == "--no-images") o.write_images = false;
        else if (a == "--quiet") o.quiet = true;
        else throw std::runtime_error("Unknown option '" + a + "'");
    }
    if (o.lanes < 1 || o.lanes > 16) throw std::runtime_error("--lanes must be in 1..16");
    if (o.seed_stride < 1) throw std::runtime_error("--seed-stride must be >= 1");
    if (o.ranks < 1 || o.ranks > 64 || o.world < 1 || o.rank < 0 || o.rank >= o.world) throw std::runtime_error("bad --ranks / --rank / --world");
    if (o.world > 1 && o.id_file.empty()) throw std::runtime_error("--world needs --id-file");
    if ((o.ranks > 1 || o.world > 1) && o.batch % (o.ranks > 1 ? o.ranks : o.world) != 0) throw std::runtime_error("--batch must divide by the number of ranks");
    if (o.K_light <= 0) o.K_light = (int)(0.2 * o.K);   // NUM_SUBSPACE_LIGHTSOURCE (optixPathTracer.h:32)
    return !(o.scene.empty() && o.cache.empty());
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// --------------------------------------------------------------------------------------------------------------
// Application state: what the reference keeps in file-scope globals (optixPathTracer.cpp:66-90)
// --------------------------------------------------------------------------------------------------------------
struct App {
    Options     opt;
    HostScene   scene;
    const HostScene* scene_ref = nullptr;   // lanes share the host copy of the scene
    spc_context* ctx = nullptr;
    spc_params  params;          // MyParams
    int         n_lvc = 0;
    std::vector<spc_tree_node> eye_tree, light_tree;
    // frame pipelining: the light trace of frame f+1 runs on a side stream under the eye pass of frame f
    cudaStream_t main_stream = nullptr, side_stream = nullptr;
    spc_vertex*  lvc[2] = {nullptr, nullptr};
    uint8_t*     valid[2] = {nullptr, nullptr};
    cudaEvent_t  ev_lt[2] = {nullptr, nullptr}, ev_eye[2] = {nullptr, nullptr};
    int          cur = 0;
    bool         primed = false;
    double       t_pretrace = 0, t_trees = 0, t_qgamma = 0;
    int          train_paths = 0;
    std::vector<float> loss;
    float*       gamma_dev = nullptr;   // trained Gamma (E), owned by the library unless loaded from a file

    template <class T> T* dalloc(size_t count) {
        void* p = nullptr;
        SPC_CHECK(spc_device_alloc(ctx, count * sizeof(T), &p));
        return static_cast<T*>(p);
    }

    // initLaunchParams + lt_params_setup + preTracer_params_setup + handleCameraUpdate
    void init_launch_params() {
        memset(&params, 0, sizeof(params));
        const size_t P = (size_t)opt.width * opt.height;
        params.width = (uint32_t)opt.width;
        params.height = (uint32_t)opt.height;
        params.accum_buffer = dalloc<spc_float4>(P);
        params.frame_buffer = dalloc<uint32_t>(P);
        params.subframe_index = 0u;
        params.max_depth = opt.max_depth;
        params.miss_color = {0.1f, 0.1f, 0.1f};
        params.subspace_info.subspaceNum = opt.K;

        spc_light_trace_params& lt = params.lt;
        lt.M_per_core = opt.lt_per_core;
        lt.core_padding = opt.lt_padding;
        lt.num_core = opt.lt_cores;
        lt.M = lt.M_per_core * lt.num_core;
        lt.launch_frame = 0;
        n_lvc = lt.num_core * lt.core_padding;
        lvc[0] = dalloc<spc_vertex>(n_lvc);
        valid[0] = dalloc<uint8_t>(n_lvc);
        lt.ans = lvc[0];
        lt.validState = valid[0];

        spc_pretrace_params& pr = params.pre_tracer;
        pr.num_core = opt.pre_cores;
        pr.padding = opt.pre_padding;
        pr.iteration = 0;
        pr.paths = dalloc<spc_train_path>(pr.num_core);
        pr.conns = dalloc<spc_train_conn>((size_t)pr.num_core * pr.padding);

        params.sky.valid = 0;   // env_params_setup with no environment file (:422-427)
        handle_camera_update();
    }

    void handle_camera_update() {
        float U[3], V[3], W[3];
        scene_ref->camera_frame(opt.width, opt.height, U, V, W);
        params.eye = {scene_ref->eye[0], scene_ref->eye[1], scene_ref->eye[2]};
        params.U = {U[0], U[1], U[2]};
        params.V = {V[0], V[1], V[2]};
        params.W = {W[0], W[1], W[2]};
    }

    void launch_light_trace() {
        params.lt.launch_frame += lt_stride;
        SPC_CHECK(spc_set_params(ctx, &params));
        SPC_CHECK(spc_launch_named(ctx, "light trace", params.lt.num_core, 1));
    }

    void launch_lvc_trace() {
        launch_light_trace();
        SPC_CHECK(spc_lvc_process(ctx, params.lt.ans, params.lt.validState, n_lvc, &params.sampler));
    }

    int launch_pretrace() {
        params.pre_tracer.iteration += pretrace_stride;
        SPC_CHECK(spc_set_params(ctx, &params));
        SPC_CHECK(spc_launch_named(ctx, "pretrace", params.pre_tracer.num_core, 1));
        int valid_samples = 0;
        SPC_CHECK(spc_valid_sample_gather(ctx, static_cast<const spc_train_path*>(params.pre_tracer.paths), params.pre_tracer.num_core,
                                          static_cast<const spc_train_conn*>(params.pre_tracer.conns),
                                          params.pre_tracer.num_core * params.pre_tracer.padding, &valid_samples));
        return valid_samples;
    }

    std::vector<spc_tree_node> build_tree(bool eye_side, int subspaces) {
        int n = 0;
        SPC_CHECK(spc_get_tree_points(ctx, eye_side ? 1 : 0, opt.tree_samples, nullptr, 0, &n));
        std::vector<spc_divide_weight> pts((size_t)(n > 0 ? n : 1));
        SPC_CHECK(spc_get_tree_points(ctx, eye_side ? 1 : 0, opt.tree_samples, pts.data(), (int)pts.size(), &n));
        std::vector<spc_tree_node> nodes((size_t)1 << 16);
        int max_label = 0;
        for (;;) {
            const int count = spc_build_tree(pts.data(), n, subspaces, 0, nodes.data(), (int)nodes.size(), &max_label);
            if (count < 0) throw std::runtime_error(std::string("spc_build_tree: ") + spc_last_error());
            if ((size_t)count <= nodes.size()) {
                nodes.resize((size_t)count);
                break;
            }
            nodes.resize((size_t)count);
        }
        if (!opt.quiet) printf("class tree building complete:size %zu and max-label %d\n", nodes.size(), max_label);
        return nodes;
    }

    // multi-GPU shard plan (the same arithmetic as spcbpt-optix7_b200/parallel.py shard_plan)
    int  pretrace_stride = 1, lt_stride = 1;

    // preprocessing() (optixPathTracer.cpp:552-608).  With world > 1 this rank traces its shard of the training paths and of the Q
    // launches; the library all-reduces the statistics over NCCL (csrc/comm.cu) and rank 0's trees are broadcast.
    void preprocessing() {
        const int rank = opt.rank, world = opt.world;
        const int local_samples = (opt.train_samples + world - 1) / world, local_q = (opt.q_samples + world - 1) / world, local_batch = opt.batch / world;
        if (world > 1) {
            pretrace_stride = lt_stride = world;
            params.pre_tracer.iteration = rank + 1 - world;
            params.lt.launch_frame = rank + 1 - world;
        }
        SPC_CHECK(spc_set_option(ctx, "train_reserve_paths", local_samples));   // size the training set once instead of doubling up to it
        double t0 = now_s();
        int current = 0;
        while (current < local_samples) {
            const int got = launch_pretrace();
            current += got;
            if (got == 0 && params.pre_tracer.iteration > 64 && current == 0) throw std::runtime_error("pretrace finds no valid path: is any light visible from the camera's paths?");
        }
        train_paths = current;
        SPC_CHECK(spc_synchronize(ctx));
        double t1 = now_s();
        t_pretrace = t1 - t0;

        SPC_CHECK(spc_sample_reweight(ctx));
        bool installed = false;
        if (rank == 0) {
            if (opt.host_trees) {
                eye_tree = build_tree(true, opt.K);
                light_tree = build_tree(false, opt.K - opt.K_light);
            } else {
                // device-side build: the weighted points never leave HBM; same trees bit for bit (csrc/tree_build.cu)
                for (int eye_side = 1; eye_side >= 0; eye_side--) {
                    std::vector<spc_tree_node>& t = eye_side ? eye_tree : light_tree;
                    spc_tree_node** dev = eye_side ? &params.subspace_info.eye_tree : &params.subspace_info.light_tree;
                    int n = 0;
                    SPC_CHECK(spc_build_tree_from_training_set(ctx, eye_side, opt.tree_samples, eye_side ? opt.K : opt.K - opt.K_light, 0, dev, &n, nullptr, 0));
                    t.resize((size_t)n);
                    SPC_CHECK(spc_download(ctx, t.data(), *dev, (size_t)n * sizeof(spc_tree_node)));
                    if (!opt.quiet) printf("class tree building complete:size %d (device)\n", n);
                }
                installed = true;
            }
        }
        if (world > 1)
            for (std::vector<spc_tree_node>* t : {&eye_tree, &light_tree}) {
                int n = (int)t->size();
                SPC_CHECK(spc_comm_bcast_host(ctx, &n, sizeof(n), 0));
                t->resize((size_t)n);
                SPC_CHECK(spc_comm_bcast_host(ctx, t->data(), (size_t)n * sizeof(spc_tree_node), 0));
            }
        if (!installed) {
            SPC_CHECK(spc_tree_to_device(ctx, 1, eye_tree.data(), (int)eye_tree.size(), &params.subspace_info.eye_tree));
            SPC_CHECK(spc_tree_to_device(ctx, 0, light_tree.data(), (int)light_tree.size(), &params.subspace_info.light_tree));
        }
        double t2 = now_s();
        t_trees = t2 - t1;

        int acc = 0;
        bool first = true;
        float* Q = nullptr;
        while (acc < local_q) {
            launch_light_trace();
            int cumulative = 0;
            SPC_CHECK(spc_preprocess_getQ(ctx, params.lt.ans, params.lt.validState, n_lvc, first ? 1 : 0, &Q, &cumulative));
            first = false;
            acc += cumulative;   // sic: the reference adds the running total each time (:590, device_thrust.cu:408)
            if (cumulative == 0) throw std::runtime_error("light trace produced no paths");
        }
        SPC_CHECK(spc_allreduce_training_stats(ctx));   // no-op on one GPU
        SPC_CHECK(spc_Q_zero_handle(ctx));
        SPC_CHECK(spc_node_label(ctx, params.subspace_info.eye_tree, params.subspace_info.light_tree));

        const int usable = current < local_samples ? current : local_samples;
        int n_train = (usable / local_batch) * local_batch;
        if (world > 1) SPC_CHECK(spc_comm_allreduce_host(ctx, &n_train, 1, SPC_COMM_I32, SPC_COMM_MIN));   // same number of Adam steps everywhere
        SPC_CHECK(spc_build_optimal_E_train_data(ctx, n_train));
        float* gamma = nullptr;
        SPC_CHECK(spc_preprocess_getGamma(ctx, &gamma));
        loss.assign(4096, 0.0f);
        int n_batches = 0;
        SPC_CHECK(spc_train_optimal_E(ctx, local_batch, opt.epochs, opt.lr, &gamma, loss.data(), (int)loss.size(), &n_batches));
        loss.resize((size_t)(n_batches < (int)loss.size() ? n_batches : (int)loss.size()));
        params.subspace_info.Q = Q;
        gamma_dev = gamma;
        SPC_CHECK(spc_Gamma2CMFGamma(ctx, gamma, &params.subspace_info.CMFGamma));
        SPC_CHECK(spc_synchronize(ctx));
        t_qgamma = now_s() - t2;
        if (world > 1) {   // the render loop's light-trace frames: disjoint from the training frames and from the other ranks'
            lt_stride = 1;
            params.lt.launch_frame = 1000003 * (opt.tiles ? 1 : rank + 1);   // tile partition: every rank traces the same light paths
        }
    }

    // trained state to / from the reference's debug text files (host/train_state.hpp)
    void save_state(const std::string& prefix) {
        spchost::TrainState st;
        st.eye_tree = eye_tree;
        st.light_tree = light_tree;
        st.Q.resize((size_t)opt.K);
        st.gamma.resize((size_t)opt.K * opt.K);
        SPC_CHECK(spc_download(ctx, st.Q.data(), params.subspace_info.Q, st.Q.size() * sizeof(float)));
        SPC_CHECK(spc_download(ctx, st.gamma.data(), gamma_dev, st.gamma.size() * sizeof(float)));
        if (!spchost::save_train_state(prefix, st)) throw std::runtime_error("cannot write training state " + prefix + "*.txt");
    }

    void load_state(const std::string& prefix) {
        spchost::TrainState st;
        std::string err;
        if (!spchost::load_train_state(prefix, opt.K, st, err)) throw std::runtime_error(err);
        eye_tree = st.eye_tree;
        light_tree = st.light_tree;
        SPC_CHECK(spc_tree_to_device(ctx, 1, eye_tree.data(), (int)eye_tree.size(), &params.subspace_info.eye_tree));
        SPC_CHECK(spc_tree_to_device(ctx, 0, light_tree.data(), (int)light_tree.size(), &params.subspace_info.light_tree));
        float* q = dalloc<float>(st.Q.size());
        gamma_dev = dalloc<float>(st.gamma.size());
        SPC_CHECK(spc_upload(ctx, q, st.Q.data(), st.Q.size() * sizeof(float)));
        SPC_CHECK(spc_upload(ctx, gamma_dev, st.gamma.data(), st.gamma.size() * sizeof(float)));
        params.subspace_info.Q = q;
        SPC_CHECK(spc_Gamma2CMFGamma(ctx, gamma_dev, &params.subspace_info.CMFGamma));
        // sample partition: every rank needs light paths of its own (as after preprocessing()); tile partition: the same everywhere
        if (opt.world > 1 && !opt.tiles) params.lt.launch_frame = 1000003 * (opt.rank + 1);
    }

    // launchSubframe (:609-635)
    void launch_subframe() {
        SPC_CHECK(spc_set_params(ctx, &params));
        SPC_CHECK(spc_launch_named(ctx, opt.alg.c_str(), opt.width, opt.height));
    }

    void enable_pipelining() {
        int lo = 0, hi = 0;
        CUDA_CHECK(cudaSetDevice(opt.device));
        CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_CHECK(cudaStreamCreateWithPriority(&main_stream, cudaStreamNonBlocking, lo));
        CUDA_CHECK(cudaStreamCreateWithPriority(&side_stream, cudaStreamNonBlocking, hi));
        lvc[1] = dalloc<spc_vertex>(n_lvc);
        valid[1] = dalloc<uint8_t>(n_lvc);
        for (int k = 0; k < 2; k++) {
            CUDA_CHECK(cudaEventCreateWithFlags(&ev_lt[k], cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&ev_eye[k], cudaEventDisableTiming));
        }
        SPC_CHECK(spc_synchronize(ctx));
        SPC_CHECK(spc_set_stream(ctx, main_stream));
    }

    void trace_into(int k) {
        params.lt.ans = lvc[k];
        params.lt.validState = valid[k];
        SPC_CHECK(spc_set_stream(ctx, side_stream));
        CUDA_CHECK(cudaStreamWaitEvent(side_stream, ev_eye[k], 0));   // the eye pass that sampled this half is done
        launch_light_trace();
        CUDA_CHECK(cudaEventRecord(ev_lt[k], side_stream));
        SPC_CHECK(spc_set_stream(ctx, main_stream));
    }

    // ---- frame lanes: this context renders the global subframes lane, lane + n_lanes, ... on its own stream ----------------------
    int          lane = 0, n_lanes = 1, lt_base = 0;
    cudaStream_t lane_stream = nullptr;

    // `owner`: a context on the same device that already holds the scene (frame lanes share lane 0's: spc_scene_share)
    void create_and_upload(double* upload_s, spc_context* owner = nullptr) {
        SPC_CHECK(spc_create(opt.device, opt.K, opt.K_light, opt.connections, &ctx));
        if (owner) {
            SPC_CHECK(spc_scene_share(ctx, owner));
            for (const auto& kv : opt.options) SPC_CHECK(spc_set_option(ctx, kv.first.c_str(), kv.second));
            if (opt.tiles && opt.world > 1) SPC_CHECK(spc_set_tile_partition(ctx, opt.rank, opt.world));
            return;
        }
        std::vector<spc_mesh> meshes;
        std::vector<spc_texture> textures;
        scene_ref->abi_views(meshes, textures);
        const double t0 = now_s();
        SPC_CHECK(spc_scene_upload(ctx, meshes.data(), (int)meshes.size(), scene_ref->materials.data(), (int)scene_ref->materials.size(), scene_ref->lights.data(),
                                   (int)scene_ref->lights.size(), textures.data(), (int)textures.size()));
        SPC_CHECK(spc_synchronize(ctx));
        if (upload_s) *upload_s = now_s() - t0;
        for (const auto& kv : opt.options) SPC_CHECK(spc_set_option(ctx, kv.first.c_str(), kv.second));
        if (opt.tiles && opt.world > 1) SPC_CHECK(spc_set_tile_partition(ctx, opt.rank, opt.world));
    }

    void become_lane(int k, int n, const App& trained) {
        lane = k;
        n_lanes = n;
        params.subspace_info = trained.params.subspace_info;   // trees, Q, CMFGamma live on the same device: shared read-only
        lt_base = trained.params.lt.launch_frame;
        CUDA_CHECK(cudaSetDevice(opt.device));
        CUDA_CHECK(cudaStreamCreateWithFlags(&lane_stream, cudaStreamNonBlocking));
        SPC_CHECK(spc_synchronize(ctx));
        SPC_CHECK(spc_set_stream(ctx, lane_stream));
        SPC_CHECK(spc_set_seed_mapping(ctx, (uint32_t)k * opt.seed_stride + opt.seed_offset, (uint32_t)n * opt.seed_stride));
        if (n > 1) SPC_CHECK(spc_set_trace_blocks(ctx, 7));   // leave room for the other lanes' small kernels (renderer.py LANE_TRACE_BLOCKS)
    }

    void render_lane_frames(int n_frames) {
        for (int f = lane; f < n_frames; f += n_lanes) {
            params.lt.launch_frame = lt_base + f;            // launch_light_trace adds 1: the index the sequential loop uses for frame f
            params.subframe_index = (uint32_t)(f / n_lanes);   // local count -> running-mean weight; the seed uses f (seed mapping)
            render_frame();
        }
        SPC_CHECK(spc_synchronize(ctx));
    }

    // one iteration of the render loop (main :796-820): light trace + LVC_Process + eye pass, or one pt subframe
    void render_frame() {
        if (opt.alg != "SPCBPT_eye") {
            launch_subframe();
        } else if (!main_stream) {
            launch_lvc_trace();
            launch_subframe();
        } else {
            if (!primed) {
                trace_into(cur);
                primed = true;
            }
            CUDA_CHECK(cudaStreamWaitEvent(main_stream, ev_lt[cur], 0));
            trace_into(1 - cur);   // next frame's light paths, under this frame's eye pass
            SPC_CHECK(spc_lvc_process(ctx, lvc[cur], valid[cur], n_lvc, &params.sampler));
            launch_subframe();
            CUDA_CHECK(cudaEventRecord(ev_eye[cur], main_stream));
            cur = 1 - cur;
        }
        ++params.subframe_index;
    }
};

HostScene load_scene(const Options& opt) {
    HostScene hs;
    std::string err;
    if (!opt.cache.empty()) {
        if (!spchost::load_scene_cache(opt.cache, hs, err)) throw std::runtime_error(err);
        return hs;
    }
    spchost::SceneFile sf;
    if (!spchost::load_scene_file(opt.scene, opt.data_root, sf, err)) throw std::runtime_error(err);
    if (!spchost::build_host_scene(sf, opt.K_light, hs)) throw std::runtime_error("scene conversion failed");
    return hs;
}

}  // namespace

int main(int argc, char** argv) {
    App app;
    try {
        if (!parse_args(argc, argv, app.opt)) {
            usage(argv[0]);
            return argc > 1 ? 0 : 1;
        }
        if (app.opt.ranks > 1) {
            // one process per GPU: fork the ranks before anything touches CUDA; rank r drives device r
            Options& o = app.opt;
            o.world = o.ranks;
            o.id_file = "/tmp/spcbpt_nccl_id_" + std::to_string((long)getpid());
            remove(o.id_file.c_str());
            std::vector<pid_t> kids;
            int my_rank = -1;
            for (int r = 0; r < o.ranks && my_rank < 0; r++) {
                const pid_t pid = fork();
                if (pid < 0) throw std::runtime_error("fork failed");
                if (pid == 0) {
                    my_rank = r;
                    prctl(PR_SET_PDEATHSIG, SIGKILL);   // a rank never outlives the launcher (it may be sitting in a collective)
                } else kids.push_back(pid);
            }
            if (my_rank < 0) {
                // Wait for the ranks.  Once one has failed the others get 30 s to finish on their own: a rank waiting in a collective
                // for a peer that is gone would otherwise keep the launcher (and its GPU) forever.
                int worst = 0;
                double failed_at = -1.0;
                while (!kids.empty()) {
                    int st = 0;
                    const pid_t pid = waitpid(-1, &st, failed_at < 0 ? 0 : WNOHANG);
                    if (pid > 0) {
                        kids.erase(std::remove(kids.begin(), kids.end(), pid), kids.end());
                        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) {
                            worst = 1;
                            if (failed_at < 0) failed_at = now_s();
                        }
                    } else if (pid < 0 && errno != EINTR) {
                        break;                              // no child left to wait for
                    } else if (failed_at >= 0 && now_s() - failed_at > 30.0) {
                        fprintf(stderr, "spcbpt_render: a rank failed; stopping the %zu rank(s) still running\n", kids.size());
                        for (pid_t k : kids) kill(k, SIGKILL);
                        failed_at = -1.0;                   // reap them with blocking waits
                    } else if (pid == 0) {
                        std::this_thread::sleep_for(std::chrono::milliseconds(100));
                    }
                }
                remove(o.id_file.c_str());
                return worst;
            }
            o.rank = my_rank;
            o.device = my_rank;
            o.ranks = 1;
            if (my_rank != 0) o.quiet = true;
        }
        if (app.opt.world > 1 && !app.opt.tiles) {   // sample partition across ranks, composed with the frame lanes below
            app.opt.seed_offset = (unsigned)app.opt.rank;
            app.opt.seed_stride = (unsigned)app.opt.world;
        }
        const Options& opt = app.opt;
        double t0 = now_s();
        app.scene = load_scene(opt);
        const double t_load = now_s() - t0;
        if (!opt.quiet) {
            for (const auto& w : app.scene.warnings) fprintf(stderr, "warning: %s\n", w.c_str());
            printf("scene: %zu meshes, %zu triangles, %zu materials, %zu lights, %zu textures (%.2f s)\n", app.scene.meshes.size(), app.scene.n_triangles(),
                   app.scene.materials.size(), app.scene.lights.size(), app.scene.textures.size(), t_load);
            printf("scene aabb is %f %f %f\n", 0.5f * (app.scene.aabb_min[0] + app.scene.aabb_max[0]), 0.5f * (app.scene.aabb_min[1] + app.scene.aabb_max[1]),
                   0.5f * (app.scene.aabb_min[2] + app.scene.aabb_max[2]));
        }
        if (!opt.save_cache.empty() && !spchost::save_scene_cache(opt.save_cache, app.scene)) throw std::runtime_error("cannot write " + opt.save_cache);
        if (!opt.render) return 0;
        if (app.scene.lights.empty()) throw std::runtime_error("the scene has no quad light: the SPCBPT path needs at least one");

        // ---- Scene::finalize(): context + upload + BVH ------------------------------------------------------------
        app.scene_ref = &app.scene;
        double t_upload = 0, t_comm = 0;
        const double t_e2e0 = now_s();   // end-to-end clock: scene upload + BVH -> training -> frames -> accumulation buffer on the host
        app.create_and_upload(&t_upload);
        spc_bvh_stats bs;
        SPC_CHECK(spc_bvh_stats_get(app.ctx, &bs));
        if (!opt.quiet) printf("bvh: %u triangles, %u 8-wide nodes, depth %u, SAH %.2f, device build %.2f ms (upload + build %.1f ms)\n", bs.n_triangles, bs.n_nodes, bs.max_depth,
                               bs.sah_cost, bs.build_ms, t_upload * 1e3);
        if (opt.seed_offset || opt.seed_stride != 1) SPC_CHECK(spc_set_seed_mapping(app.ctx, opt.seed_offset, opt.seed_stride));
        if (opt.world > 1) {
            // NCCL rendezvous through a file: rank 0 creates the unique id, the others wait for it
            unsigned char id[SPC_COMM_ID_BYTES];
            if (opt.rank == 0) {
                if (spc_comm_unique_id(id) != SPC_OK) throw std::runtime_error(std::string("spc_comm_unique_id: ") + spc_last_error());
                const std::string tmp = opt.id_file + ".tmp";
                std::ofstream(tmp, std::ios::binary).write(reinterpret_cast<const char*>(id), sizeof(id));
                if (rename(tmp.c_str(), opt.id_file.c_str()) != 0) throw std::runtime_error("cannot write " + opt.id_file);
            } else {
                bool got = false;
                for (int tries = 0; tries < 1200 && !got; tries++) {
                    std::ifstream f(opt.id_file, std::ios::binary);
                    got = f && f.read(reinterpret_cast<char*>(id), sizeof(id)) && f.gcount() == (std::streamsize)sizeof(id);
                    if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(100));
                }
                if (!got) throw std::runtime_error("timed out waiting for the NCCL id in " + opt.id_file);
            }
            const double tc0 = now_s();
            SPC_CHECK(spc_comm_init(app.ctx, opt.rank, opt.world, id));
            t_comm = now_s() - tc0;   // NCCL bootstrap + communicator + warm-up collective: tens of seconds for 8 ranks on a cold node
        }

        app.init_launch_params();
        const bool lanes_on = opt.lanes > 1;
        if (opt.alg == "SPCBPT_eye") {
            if (!opt.quiet) printf("BDPTVertex Size %zu\n", sizeof(spc_vertex));
            if (!opt.load_state.empty()) app.load_state(opt.load_state);
            else app.preprocessing();
            if (!opt.save_state.empty()) app.save_state(opt.save_state);
            if (!opt.quiet && opt.load_state.empty())
                printf("preprocessing: %d training paths in %.3f s, trees (%zu + %zu nodes) %.3f s, Q + Gamma training %.3f s, loss %.6f -> %.6f\n", app.train_paths, app.t_pretrace,
                       app.eye_tree.size(), app.light_tree.size(), app.t_trees, app.t_qgamma, app.loss.empty() ? 0.f : app.loss.front(), app.loss.empty() ? 0.f : app.loss.back());
            if (opt.pipeline && !lanes_on) app.enable_pipelining();
        } else if (opt.alg != "pt") {
            throw std::runtime_error("unknown integrator '" + opt.alg + "' (pt | SPCBPT_eye)");
        }

        // ---- frame lanes: more contexts on the same device sharing the trained state ------------------------------------
        std::vector<std::unique_ptr<App>> extra;
        std::vector<App*> lanes{&app};
        if (lanes_on) {
            for (int k = 1; k < opt.lanes; k++) {
                extra.emplace_back(new App());
                App& l = *extra.back();
                l.opt = opt;
                l.scene_ref = &app.scene;
                // lanes borrow lane 0's scene and BVH (spc_scene_share).  Verified on one GPU; multi-rank runs keep a copy per lane, the
                // configuration that was measured on 2 and 8 GPUs before the round's GPU budget ran out (profiles/r2_summary.md section 5)
                l.create_and_upload(nullptr, opt.world == 1 ? app.ctx : nullptr);
                l.init_launch_params();
                lanes.push_back(&l);
            }
            for (int k = 0; k < opt.lanes; k++) lanes[k]->become_lane(k, opt.lanes, app);
        }

        // ---- render loop ----------------------------------------------------------------------------------------------
        SPC_CHECK(spc_synchronize(app.ctx));
        SPC_CHECK(spc_comm_barrier(app.ctx));
        int64_t launches = 0;
        for (App* l : lanes) launches -= spc_launch_count(l->ctx);
        t0 = now_s();
        if (!lanes_on) {
            for (int f = 0; f < opt.frames; f++) app.render_frame();
            SPC_CHECK(spc_synchronize(app.ctx));
            if (app.side_stream) CUDA_CHECK(cudaStreamSynchronize(app.side_stream));
        } else {
            std::vector<std::thread> threads;
            std::vector<std::exception_ptr> errors(lanes.size());
            for (size_t k = 0; k < lanes.size(); k++)
                threads.emplace_back([&, k] {
                    try {
                        lanes[k]->render_lane_frames(opt.frames);
                    } catch (...) {
                        errors[k] = std::current_exception();
                    }
                });
            for (auto& t : threads) t.join();
            for (auto& e : errors)
                if (e) std::rethrow_exception(e);
        }
        double t_render = now_s() - t0;
        if (opt.world > 1) SPC_CHECK(spc_comm_allreduce_host(app.ctx, &t_render, 1, SPC_COMM_F64, SPC_COMM_MAX));   // the slowest rank
        for (App* l : lanes) launches += spc_launch_count(l->ctx);

        if (lanes_on) {   // read-out: merge the running means, weights = share of the subframes each lane rendered
            std::vector<const spc_float4*> bufs;
            std::vector<float> weights;
            for (int k = 0; k < opt.lanes; k++) {
                const int n_k = (opt.frames - k + opt.lanes - 1) / opt.lanes;
                if (n_k <= 0) continue;
                bufs.push_back(lanes[k]->params.accum_buffer);
                weights.push_back((float)n_k / (float)opt.frames);
            }
            spc_float4* merged = app.dalloc<spc_float4>((size_t)opt.width * opt.height);
            SPC_CHECK(spc_merge_accum(app.ctx, bufs.data(), weights.data(), (int)bufs.size(), opt.width * opt.height, merged, app.params.frame_buffer));
            app.params.accum_buffer = merged;
        }

        if (opt.world > 1) {
            // read-out over NCCL: every rank rendered opt.frames subframes -> the image is the mean of the per-rank running means
            // (tile partition: every pixel is non-zero on exactly one rank -> weight 1)
            SPC_CHECK(spc_reduce_accum(app.ctx, app.params.accum_buffer, opt.width * opt.height, opt.tiles ? 1.0f : 1.0f / (float)opt.world, 0));
            SPC_CHECK(spc_synchronize(app.ctx));
            if (opt.rank == 0) {   // display transform of the merged image
                const spc_float4* one = app.params.accum_buffer;
                const float w1 = 1.0f;
                SPC_CHECK(spc_merge_accum(app.ctx, &one, &w1, 1, opt.width * opt.height, nullptr, app.params.frame_buffer));
            } else {
                for (auto& l : extra) spc_destroy(l->ctx);
                spc_destroy(app.ctx);
                return 0;
            }
        }
        const size_t P = (size_t)opt.width * opt.height;
        std::vector<float> accum(P * 4);
        std::vector<uint32_t> frame(P);
        SPC_CHECK(spc_download(app.ctx, accum.data(), app.params.accum_buffer, accum.size() * sizeof(float)));
        SPC_CHECK(spc_download(app.ctx, frame.data(), app.params.frame_buffer, frame.size() * sizeof(uint32_t)));
        const double t_e2e = now_s() - t_e2e0;
        double mean = 0;
        for (size_t i = 0; i < P; i++) mean += accum[4 * i] + accum[4 * i + 1] + accum[4 * i + 2];
        mean /= (double)(3 * P);
        if (opt.write_images) {
            if (!spchost::write_ppm_from_uchar4(opt.out + ".ppm", frame.data(), opt.width, opt.height)) throw std::runtime_error("cannot write " + opt.out + ".ppm");
            if (!spchost::write_png_from_uchar4(opt.out + ".png", frame.data(), opt.width, opt.height)) throw std::runtime_error("cannot write " + opt.out + ".png");
            if (!spchost::write_pfm_from_float4(opt.out + ".pfm", accum.data(), opt.width, opt.height)) throw std::runtime_error("cannot write " + opt.out + ".pfm");
        }
        const int sample_ranks = opt.tiles ? 1 : opt.world;   // tile partition: the ranks share the subframes instead of adding their own
        printf("{\"alg\": \"%s\", \"width\": %d, \"height\": %d, \"frames\": %d, \"ranks\": %d, \"triangles\": %zu, \"K\": %d, \"lanes\": %d, \"pipelined\": %s, \"render_s\": %.6f, \"ms_per_frame\": %.4f, "
               "\"samples_per_s\": %.1f, \"kernel_launches\": %lld, \"train_paths\": %d, \"pretrace_s\": %.4f, \"trees_s\": %.4f, \"q_gamma_s\": %.4f, \"upload_s\": %.4f, \"comm_init_s\": %.4f, \"e2e_s\": %.4f, "
               "\"e2e_samples_per_s\": %.1f, \"h2d_bytes\": %zu, \"d2h_bytes\": %zu, \"image_mean\": %.9g}\n",
               opt.alg.c_str(), opt.width, opt.height, opt.frames, opt.world, app.scene.n_triangles(), opt.K, opt.lanes, app.main_stream ? "true" : "false", t_render, t_render / opt.frames * 1e3,
               (double)P * opt.frames * sample_ranks / t_render, (long long)launches, app.train_paths, app.t_pretrace, app.t_trees, app.t_qgamma, t_upload, t_comm, t_e2e,
               (double)P * opt.frames * sample_ranks / t_e2e, app.scene.upload_bytes(), P * 20, mean);
        for (auto& l : extra) spc_destroy(l->ctx);
        spc_destroy(app.ctx);
    } catch (std::exception& e) {
        fprintf(stderr, "Caught exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
