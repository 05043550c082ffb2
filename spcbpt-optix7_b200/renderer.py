"""Host-side mirror of the reference application's schedule (src/OptiXPathTracer/optixPathTracer.cpp) on top of the C ABI:
buffer setup (initLaunchParams :260-310, lt_params_setup :462-476, preTracer_params_setup :479-488), preprocessing()
(:552-608) and the per-frame loop (launchLVCTrace :515-522, launchSubframe :609-635).  Python here is only the
sequencer -- every step is a call into libspcbpt_b200.so; torch is used as the device allocator and for NCCL.

The C++ twin of this file is host/spcbpt_main.cpp (same calls, no Python)."""
import time

import numpy as np

from . import (LAUNCH_LIGHT_TRACE, LAUNCH_PRETRACE, LAUNCH_PT, LAUNCH_SPCBPT_EYE, PARAMS, TRAIN_CONN, TRAIN_PATH, TREE_NODE, VERTEX, Context, build_tree)


class Renderer:
    def __init__(self, scene, width, height, device=0, K=1000, K_light=0, connections=3, max_depth=0,
                 lt_num_core=1000, lt_core_padding=800, lt_M_per_core=100, pretrace_num_core=10000, pretrace_padding=10, stream=None, fast=False,
                 scene_owner=None):
        import torch
        self.torch = torch
        self.scene, self.w, self.h = scene, width, height
        self.K = K
        self.K_light = K_light if K_light else int(0.2 * K)
        self.ctx = Context(device, K=K, K_light=self.K_light, connections=connections, fast=fast)   # fast: the fast-arithmetic flavour
        if stream is not None:
            self.ctx.set_stream(stream)
        if scene_owner is not None:     # another context on this device already holds the scene: one replica per GPU (spc_scene_share)
            self.ctx.share_scene(scene_owner)
        else:
            self.ctx.upload_scene(scene)
        dev = torch.device("cuda", device)
        self.dev = dev
        P = np.zeros(1, PARAMS)
        self.P = P
        eye, U, V, W = scene.camera_frame(width, height)
        P["width"], P["height"], P["max_depth"] = width, height, max_depth
        P["eye"], P["U"], P["V"], P["W"] = eye, U, V, W
        self.accum = torch.zeros((width * height, 4), dtype=torch.float32, device=dev)
        self.frame = torch.zeros(width * height, dtype=torch.int32, device=dev)
        P["accum_buffer"], P["frame_buffer"] = self.accum.data_ptr(), self.frame.data_ptr()
        n = lt_num_core * lt_core_padding
        self.n_lvc = n
        self.lvc = torch.zeros(n * VERTEX.itemsize, dtype=torch.uint8, device=dev)
        self.valid = torch.zeros(n, dtype=torch.uint8, device=dev)
        lt = P["lt"]
        lt["num_core"], lt["core_padding"], lt["M_per_core"], lt["M"] = lt_num_core, lt_core_padding, lt_M_per_core, lt_num_core * lt_M_per_core
        lt["ans"], lt["validState"] = self.lvc.data_ptr(), self.valid.data_ptr()
        self.tp = torch.zeros(pretrace_num_core * TRAIN_PATH.itemsize, dtype=torch.uint8, device=dev)
        self.tc = torch.zeros(pretrace_num_core * pretrace_padding * TRAIN_CONN.itemsize, dtype=torch.uint8, device=dev)
        pt = P["pre_tracer"]
        pt["num_core"], pt["padding"], pt["iteration"] = pretrace_num_core, pretrace_padding, 0
        pt["paths"], pt["conns"] = self.tp.data_ptr(), self.tc.data_ptr()
        P["subspace_info"]["subspaceNum"] = K
        self.subframe = 0
        self.stats = {}
        self._pipe = None
        self.pretrace_stride = self.lt_stride = 1     # multi-GPU training shards the launches: parallel.shard_plan
        self.device_trees = True                      # classification trees built on the GPU (False: the host builder, as the reference)

    # ---- launch helpers (optixPathTracer.cpp:491-549) ------------------------------------------
    def launch_light_trace(self):
        self.P["lt"]["launch_frame"] += self.lt_stride
        self.ctx.set_params(self.P)
        self.ctx.launch(LAUNCH_LIGHT_TRACE, int(self.P["lt"]["num_core"][0]), 1)

    def launch_lvc_trace(self):
        self.launch_light_trace()
        self.P["sampler"] = self.ctx.lvc_process(self.lvc, self.valid, self.n_lvc)[0]

    def launch_pretrace(self):
        pt = self.P["pre_tracer"]
        pt["iteration"] += self.pretrace_stride
        self.ctx.set_params(self.P)
        n = int(pt["num_core"][0])
        self.ctx.launch(LAUNCH_PRETRACE, n, 1)
        return self.ctx.valid_sample_gather(self.tp, n, self.tc, n * int(pt["padding"][0]))

    def launch_subframe(self):
        self.P["subframe_index"] = self.subframe
        self.ctx.set_params(self.P)
        self.ctx.launch(LAUNCH_SPCBPT_EYE, self.w, self.h)

    # ---- preprocessing() (optixPathTracer.cpp:552-608) -----------------------------------------
    def preprocessing(self, target_samples=2000000, target_Q_samples=2000000, tree_samples=100000, batch_size=20000, epochs=1, lr=0.01,
                      plan=None, verbose=False, adam=True):
        """the subspace-training schedule.  Multi-GPU: `plan` = parallel.shard_plan(...) and the context has a communicator
        (parallel.comm_init): this rank traces its shard of the training paths and of the Q launches, the library all-reduces
        the statistics (csrc/comm.cu), rank 0 builds the trees."""
        ctx, K = self.ctx, self.K
        rank, world = (plan["rank"], plan["world"]) if plan else (0, 1)
        local_samples = plan["local_samples"] if plan else target_samples
        local_Q = plan["local_Q_samples"] if plan else target_Q_samples
        local_batch = plan["local_batch"] if plan else batch_size
        pt, lt = self.P["pre_tracer"], self.P["lt"]
        if plan:
            self.pretrace_stride, self.lt_stride = plan["iteration_stride"], plan["lt_frame_stride"]
            pt["iteration"] = plan["first_iteration"] - self.pretrace_stride
            lt["launch_frame"] = plan["first_lt_frame"] - self.lt_stride
        ctx.set_option("train_reserve_paths", int(local_samples))   # size the training set once instead of doubling up to it
        t0 = time.perf_counter()
        n = 0
        while n < local_samples:
            n += self.launch_pretrace()
        ctx.synchronize()
        t1 = time.perf_counter()
        ctx.sample_reweight()
        si = self.P["subspace_info"]
        eye_tree = light_tree = None
        if rank == 0:    # rank 0 builds the trees from its shard of the paths
            if self.device_trees:    # on the device: the weighted points never leave HBM (csrc/tree_build.cu; same trees bit for bit)
                si["eye_tree"], eye_tree = ctx.build_tree_from_training_set(True, tree_samples, K, 0)
                si["light_tree"], light_tree = ctx.build_tree_from_training_set(False, tree_samples, K - self.K_light, 0)
            else:                    # the reference's way: points to the host, host builder, tree back to the device
                eye_tree, _ = build_tree(ctx.get_tree_points(True, tree_samples), K, 0)
                light_tree, _ = build_tree(ctx.get_tree_points(False, tree_samples), K - self.K_light, 0)
        if world > 1:
            eye_tree, light_tree = (ctx.comm_bcast_array(t, TREE_NODE, 0) for t in (eye_tree, light_tree))
        self.eye_tree, self.light_tree = eye_tree, light_tree
        if not (rank == 0 and self.device_trees):
            si["eye_tree"] = ctx.tree_to_device(True, eye_tree)
            si["light_tree"] = ctx.tree_to_device(False, light_tree)
        t2 = time.perf_counter()
        acc, first, q_dev = 0, True, 0
        while acc < local_Q:
            self.launch_light_trace()
            q_dev, cum = ctx.preprocess_getQ(self.lvc, self.valid, self.n_lvc, reset=first)
            first = False
            acc += cum      # sic: the reference adds the CUMULATIVE count each time (optixPathTracer.cpp:590, device_thrust.cu:408)
        ctx.allreduce_training_stats()      # no-op on one GPU
        ctx.Q_zero_handle()
        ctx.node_label(int(si["eye_tree"][0]), int(si["light_tree"][0]))
        n_train = (min(n, local_samples) // local_batch) * local_batch
        if world > 1:       # every rank must run the same number of Adam steps
            n_train = int(ctx.comm_allreduce_host(np.array([n_train], np.int32), "min")[0])
        ctx.build_optimal_E_train_data(n_train)
        g_dev = ctx.preprocess_getGamma()
        loss = np.zeros(0, np.float32)
        if adam:    # (adam=False keeps the row-normalised histogram Gamma: the study in tests/quick_adam_vs_histogram.py)
            g_dev, loss = ctx.train_optimal_E(local_batch, epochs, lr)
        si["Q"] = q_dev
        self.gamma_dev = g_dev
        si["CMFGamma"] = ctx.Gamma2CMFGamma(g_dev)
        ctx.synchronize()
        t3 = time.perf_counter()
        if plan:
            self.lt_stride = 1
            lt["launch_frame"] = plan["render_lt_base"]
        self.stats.update(train_paths=n, train_paths_used=n_train, pretrace_s=t1 - t0, trees_s=t2 - t1, q_gamma_s=t3 - t2, loss_first=float(loss[0]) if len(loss) else None,
                          loss_last=float(loss[-1]) if len(loss) else None, eye_tree_nodes=int(eye_tree.shape[0]), light_tree_nodes=int(light_tree.shape[0]))
        if verbose:
            print(self.stats)
        return self.stats

    # ---- trained state on disk: the reference's debug text files (classTree::tree_load, load_Q_file, load_Gamma_file;
    # same files as host/train_state.cpp) ------------------------------------------------------------------------------
    def save_state(self, prefix):
        si = self.P["subspace_info"]
        for name, tree in (("tree_eye.txt", self.eye_tree), ("tree_light.txt", self.light_tree)):
            with open(prefix + name, "w") as f:
                for n in tree:
                    if n["leaf"]:
                        f.write("1 %d\n" % n["label"])
                    else:
                        f.write("0 %d %d %.9g %.9g %.9g %s\n" % (n["label"], n["type"], n["mid"][0], n["mid"][1], n["mid"][2], " ".join(str(int(c)) for c in n["child"])))
        np.savetxt(prefix + "Q.txt", self.ctx.download(int(si["Q"][0]), np.float32, self.K), fmt="%.9g")
        np.savetxt(prefix + "E.txt", self.ctx.download(int(self.gamma_dev), np.float32, self.K * self.K).reshape(self.K, self.K), fmt="%.9g")

    def load_state(self, prefix):
        from . import TREE_NODE
        torch = self.torch
        trees = []
        for name in ("tree_eye.txt", "tree_light.txt"):
            tok = open(prefix + name).read().split()
            nodes, i = [], 0
            while i < len(tok):
                n = np.zeros(1, TREE_NODE)[0]
                n["leaf"], n["label"] = int(tok[i]), int(tok[i + 1])
                i += 2
                if not n["leaf"]:
                    n["type"] = int(tok[i])
                    n["mid"] = [np.float32(t) for t in tok[i + 1:i + 4]]
                    n["child"] = [int(t) for t in tok[i + 4:i + 12]]
                    i += 12
                nodes.append(n)
            trees.append(np.array(nodes, TREE_NODE))
        self.eye_tree, self.light_tree = trees
        si = self.P["subspace_info"]
        si["eye_tree"] = self.ctx.tree_to_device(True, self.eye_tree)
        si["light_tree"] = self.ctx.tree_to_device(False, self.light_tree)
        self._q_t = torch.from_numpy(np.loadtxt(prefix + "Q.txt", dtype=np.float32)).to(self.dev)
        self._g_t = torch.from_numpy(np.loadtxt(prefix + "E.txt", dtype=np.float32).reshape(-1)).to(self.dev)
        assert self._q_t.numel() == self.K and self._g_t.numel() == self.K * self.K
        torch.cuda.synchronize(self.dev)
        si["Q"] = self._q_t.data_ptr()
        self.gamma_dev = self._g_t.data_ptr()
        si["CMFGamma"] = self.ctx.Gamma2CMFGamma(self.gamma_dev)

    def _as_tensor(self, dev_ptr, count):
        """wrap a device pointer owned by the context as a float32 torch tensor (for torch.distributed collectives)"""
        torch = self.torch
        iface = {"shape": (count,), "typestr": "<f4", "data": (int(dev_ptr), False), "version": 2}
        holder = type("DevArray", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=self.dev)

    # ---- the per-frame loop (optixPathTracer.cpp:791-822 without the GL display) ------------------
    def render_frame(self):
        if self._pipe is not None:
            return self._render_frame_pipelined()
        self.launch_lvc_trace()
        self.launch_subframe()
        self.subframe += 1

    def enable_pipelining(self):
        """Overlap the (latency-bound, 1000-lane) light trace of frame f+1 with the eye pass of frame f: the light trace runs on
        a side stream into the other half of a double-buffered LVC.  Same launches, same seeds, same results as render_frame()."""
        torch = self.torch
        with torch.cuda.device(self.dev):
            main = torch.cuda.Stream()
            side = torch.cuda.Stream(priority=-1)
            lvc2 = torch.zeros_like(self.lvc)
            valid2 = torch.zeros_like(self.valid)
            self._pipe = dict(main=main, side=side, lvc=[self.lvc, lvc2], valid=[self.valid, valid2], cur=0, primed=False,
                              ev_lt=[torch.cuda.Event(), torch.cuda.Event()], ev_eye=[torch.cuda.Event(), torch.cuda.Event()])
        main.wait_stream(torch.cuda.current_stream(self.dev))
        self.ctx.set_stream(main.cuda_stream)

    def _trace_into(self, k):
        p = self._pipe
        lt = self.P["lt"]
        lt["ans"], lt["validState"] = p["lvc"][k].data_ptr(), p["valid"][k].data_ptr()
        self.ctx.set_stream(p["side"].cuda_stream)
        p["side"].wait_event(p["ev_eye"][k])      # the eye pass that sampled this half must be done
        self.launch_light_trace()
        p["ev_lt"][k].record(p["side"])
        self.ctx.set_stream(p["main"].cuda_stream)

    def _render_frame_pipelined(self):
        p = self._pipe
        cur = p["cur"]
        if not p["primed"]:
            self._trace_into(cur)
            p["primed"] = True
        p["main"].wait_event(p["ev_lt"][cur])
        self._trace_into(1 - cur)                   # next frame's light paths, under this frame's eye pass
        self.P["sampler"] = self.ctx.lvc_process(p["lvc"][cur], p["valid"][cur], self.n_lvc)[0]
        self.launch_subframe()
        p["ev_eye"][cur].record(p["main"])
        p["cur"] = 1 - cur
        self.subframe += 1

    def render_frame_pt(self):
        """one subframe of the "pt" comparison integrator (Space key in the reference UI, optixPathTracer.cpp:198-208)"""
        self.P["subframe_index"] = self.subframe
        self.ctx.set_params(self.P)
        self.ctx.launch(LAUNCH_PT, self.w, self.h)
        self.subframe += 1

    def reset_accumulation(self):
        self.subframe = 0

    def image(self):
        """(H, W, 3) float32 accumulated radiance"""
        self.ctx.synchronize()
        return self.accum.cpu().numpy()[:, :3].reshape(self.h, self.w, 3)

    def frame_rgba8(self):
        self.ctx.synchronize()
        return self.frame.cpu().numpy().view(np.uint8).reshape(self.h, self.w, 4)


LANE_TRACE_BLOCKS = 7   # resident blocks per SM of the persistent trace kernels when several lanes share a GPU (host/spcbpt_main.cpp uses the same)


class LaneRenderer:
    """Frame lanes: `lanes` render contexts on one GPU, each on its own stream and driven by its own host thread, rendering
    alternate subframes (lane k draws the samples of the global subframes k, k+lanes, ... through spc_set_seed_mapping and
    keeps its own running mean).  The low-occupancy tail of one lane's eye pass (deep bounces with few live paths, each a
    latency-bound kernel) and its light trace run under the full-width head of another lane's frame.  Read-out merges the
    running means with weights n_k / n (spc_merge_accum): the same sample set as the sequential loop, summed in a different
    fp32 order.  Training happens once, on lane 0; the other lanes share its trees, Q and CMFGamma (same device)."""

    def __init__(self, scene, width, height, lanes=3, device=0, share_scene=True, **kw):
        import torch
        self.torch = torch
        self.w, self.h, self.n_lanes = width, height, lanes
        self.lanes = [Renderer(scene, width, height, device=device, **kw)]
        for _ in range(1, lanes):       # the other lanes read lane 0's scene and BVH: one replica per GPU (share_scene=False: own copies)
            self.lanes.append(Renderer(scene, width, height, device=device, scene_owner=self.lanes[0].ctx if share_scene else None, **kw))
        with torch.cuda.device(device):
            self.streams = [torch.cuda.Stream() for _ in range(lanes)]
        for k, (r, s) in enumerate(zip(self.lanes, self.streams)):
            r.ctx.synchronize()
            r.ctx.set_stream(s.cuda_stream)
            r.ctx.set_seed_mapping(k, lanes)
            if lanes > 1:
                r.ctx.set_trace_blocks(LANE_TRACE_BLOCKS)   # leave room for the other lanes' small kernels
        self.ctx = self.lanes[0].ctx
        self.frames = 0            # global subframes rendered so far
        self.lt_base = 0
        self.accum = torch.zeros((width * height, 4), dtype=torch.float32, device=self.lanes[0].dev)
        self.frame = torch.zeros(width * height, dtype=torch.int32, device=self.lanes[0].dev)

    def preprocessing(self, **kw):
        st = self.lanes[0].preprocessing(**kw)
        self.share_trained_state()
        return st

    def share_trained_state(self):
        """after lane 0 has been trained (Renderer.preprocessing / parallel.preprocess_distributed / load_state)"""
        r0 = self.lanes[0]
        for r in self.lanes[1:]:
            r.P["subspace_info"] = r0.P["subspace_info"]
            r.eye_tree, r.light_tree = r0.eye_tree, r0.light_tree
        self.lt_base = int(r0.P["lt"]["launch_frame"][0])

    def seed_mapping(self, offset, stride):
        """compose with an outer partition (multi-GPU): global sample index = (local index) * stride + offset"""
        for k, r in enumerate(self.lanes):
            r.ctx.set_seed_mapping(k * stride + offset, self.n_lanes * stride)

    def _worker(self, k, f0, f1, errors):
        try:
            r = self.lanes[k]
            first = f0 + ((k - f0) % self.n_lanes)
            for f in range(first, f1, self.n_lanes):
                r.P["lt"]["launch_frame"] = self.lt_base + f     # render_frame adds 1: frame f uses light-trace index lt_base+f+1
                r.subframe = f // self.n_lanes
                r.render_frame()
            r.ctx.synchronize()
        except Exception as ex:   # surfaced by render()
            errors.append(ex)

    def render(self, n_frames):
        """render the next n_frames global subframes (ctypes releases the GIL inside every library call)"""
        import threading
        errors = []
        f0, f1 = self.frames, self.frames + n_frames
        th = [threading.Thread(target=self._worker, args=(k, f0, f1, errors)) for k in range(self.n_lanes)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errors:
            raise errors[0]
        self.frames = f1

    def lane_counts(self):
        return [len(range(k, self.frames, self.n_lanes)) for k in range(self.n_lanes)]

    def merge(self):
        cnt = self.lane_counts()
        used = [(r.accum, c / self.frames) for r, c in zip(self.lanes, cnt) if c > 0]
        self.ctx.merge_accum([a for a, _ in used], [w for _, w in used], self.w * self.h, self.accum, self.frame)
        return self.accum

    def image(self):
        self.merge()
        self.ctx.synchronize()
        return self.accum.cpu().numpy()[:, :3].reshape(self.h, self.w, 3)

    def frame_rgba8(self):
        self.merge()
        self.ctx.synchronize()
        return self.frame.cpu().numpy().view(np.uint8).reshape(self.h, self.w, 4)

    def launch_count(self):
        return sum(r.ctx.launch_count() for r in self.lanes)
