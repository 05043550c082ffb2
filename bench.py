#!/usr/bin/env python
"""bench.py -- headline benchmark of the SPCBPT render core on B200 (contract: see the task brief).

Workload at every N (weak scaling, one replica of scene+BVH per GPU, no data-path collective):
BASELINE.json configs[1], the BVH traversal microbench -- fractal height field 708x708 quads
(1 002 542 triangles with its box and light) and three ray sets of 2^24 rays per GPU:
  A coherent pinhole primaries 4096x4096      (closest hit)
  B incoherent cosine-bounce rays from A's hits (closest hit)   <- dominant kernel, roofline
  C shadow rays from A's hits to the quad light (occlusion)
One "step" = one pass of the traversal hot path over the three batches (3 kernel launches).
metric = Mrays/s (BASELINE.json: "Mrays/s ... vs HBM roofline").

  value  rays resident in HBM when the timed region starts (CUDA events on the launching stream)
  e2e    the same three batches through the host-buffer C-ABI calls (spc_trace_batch /
         spc_occlusion_batch): pinned host rays -> H2D -> kernel -> D2H hits, all inside the timing
  roofline      kernel B: algorithmic bytes = rays * (32 + 16 + 80*n_node + 48*n_tri) with n_node, n_tri
                measured by the instrumented copy of the same kernel, / its mean launch time (events)
  cpu_baseline  oracle port (oracle/orc_scene.cpp) on all host cores over a bounded sample

`--impl reference` times the CPU implementation of the same path: the reference's own traversal is
OptiX (closed source, absent) so this arm is the oracle port (kind "port") on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "bvh_traversal_microbench: heightfield 708x708 (1002542 tris), ray sets A primary 4096^2 + B cosine-bounce + C shadow, 2^24 rays each per GPU"
RAYS_SIDE = 4096


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


T_START = time.perf_counter()


def note(msg):
    """progress line on stderr (stdout carries the one JSON line): when a run stops short, the log says where it stood"""
    sys.stderr.write("[bench %7.1f s rank %s] %s\n" % (time.perf_counter() - T_START, os.environ.get("RANK", "0"), msg))
    sys.stderr.flush()


ORIG_AFFINITY = None
BOUND_AFFINITY = None


def restore_affinity():
    """the CPU baseline legs use every host core"""
    if ORIG_AFFINITY:
        os.sched_setaffinity(0, ORIG_AFFINITY)


def rebind_affinity():
    """back onto the GPU's NUMA node after a CPU baseline leg (no-op when bind_to_gpu_numa_node did not bind)"""
    if BOUND_AFFINITY:
        try:
            os.sched_setaffinity(0, BOUND_AFFINITY)
        except OSError:
            pass


class SectionGuard:
    """Bounded time for the extra `spcbpt` section.  The headline line (traversal metric, roofline, parity, cpu_baseline) is complete
    before the section starts; if the section is still running after `seconds` -- a rank stuck in a collective, a lost GPU -- every
    rank dumps its threads' stacks to stderr and leaves with status 0, rank 0 after printing the headline line with the section
    marked as timed out.  The line is printed exactly once: by finish() on the main thread or by the timer, never both."""

    def __init__(self, seconds, rank, line):
        self.seconds, self.rank, self.line = seconds, rank, line
        self.lock = threading.Lock()
        self.done = False
        self.timer = None

    def start(self):
        if self.seconds > 0:
            # rank 0 first: it owns the line; the others follow a few seconds later so that no peer disappears under it
            self.timer = threading.Timer(self.seconds + (0 if self.rank == 0 else 5), self._fire)
            self.timer.daemon = True
            self.timer.start()
        return self

    def _fire(self):
        with self.lock:
            if self.done:
                return
            try:
                import faulthandler
                sys.stderr.write("bench.py: the SPCBPT section is still running after %d s on rank %d; thread stacks follow, the headline "
                                 "line is printed without the section\n" % (self.seconds, self.rank))
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                sys.stderr.flush()
                self._print("section timed out after %d s (thread stacks on stderr); headline unaffected" % self.seconds)
            finally:
                os._exit(0)

    def _print(self, why):
        if self.rank == 0 and self.line is not None:
            self.line["spcbpt"] = {"error": why}
            sys.stdout.write(json.dumps(self.line) + "\n")
            sys.stdout.flush()

    def abandon(self, why):
        """the section failed on this rank of a multi-rank run: print the headline (rank 0) and leave the process; never returns"""
        with self.lock:
            try:
                if not self.done:
                    self._print(why)
                sys.stderr.flush()
            finally:
                os._exit(0)

    def finish(self):
        """the section returned (or raised): from here on the main thread owns the line"""
        with self.lock:
            self.done = True
        if self.timer is not None:
            self.timer.cancel()


def bind_to_gpu_numa_node(gpu_index):
    """Keep this rank's host threads and its pinned buffers on the NUMA node its GPU hangs off (the host-buffer `e2e` leg moves 72 GB/s
    per rank through host memory; unbound ranks of a multi-GPU run land on one socket).  Best effort: returns a note for `config`."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bdf = out[-12:] if len(out) >= 12 else out          # nvidia-smi prints an 8-digit domain: keep dddd:bb:dd.f
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        global ORIG_AFFINITY, BOUND_AFFINITY
        ORIG_AFFINITY = os.sched_getaffinity(0)
        cpus &= ORIG_AFFINITY
        if len(cpus) < 8:     # a rank drives 4 frame lanes (spinning host threads) next to NCCL's proxy thread: do not squeeze them
            return "numa: node %d offers %d usable cpus, not bound" % (node, len(cpus))
        os.sched_setaffinity(0, cpus)
        BOUND_AFFINITY = cpus
        return "numa: bound to node %d (%d cpus)" % (node, len(cpus))
    except Exception as ex:   # topology files absent (VM without NUMA information): leave the affinity alone
        return "numa: not bound (%s)" % type(ex).__name__


def cpu_trace_sample(pkg, orc, scene, sets, n_sample, threads):
    """oracle port on the host cores over the first n_sample rays of each set; returns (Mrays/s, seconds, rays, results):
    the results are kept -- they are the parity check of the GPU outputs on the very rays the bench times"""
    osc = orc.Scene(pkg, scene)
    t0 = time.perf_counter()
    total, results = 0, []
    for kind, rays in sets:
        r = rays[:n_sample]
        if kind == "occlusion":
            results.append(osc.occlusion(r, threads=threads))
        else:
            results.append(osc.trace(r, threads=threads))
        total += r.shape[0]
    dt = time.perf_counter() - t0
    return total / dt / 1e6, dt, total, results


def check_parity(pkg, gpu_hits, gpu_vis, oracle_results, ns):
    """GPU results of the timed kernels against the oracle on the same rays: prim ids, t/u/v bits, visibility -- all exact"""
    out = {"rays_checked": 0, "prim_mismatch": 0, "tuv_bit_mismatch": 0, "visibility_mismatch": 0}
    for g, o in zip(gpu_hits, oracle_results[:2]):
        g = g[:ns].cpu().numpy().view(pkg.HIT).reshape(-1)
        out["rays_checked"] += int(o.shape[0])
        out["prim_mismatch"] += int((g["prim"] != o["prim"]).sum())
        for k in ("t", "u", "v"):
            out["tuv_bit_mismatch"] += int((g[k].view(np.uint32) != o[k].view(np.uint32)).sum())
    v = gpu_vis[:ns].cpu().numpy()
    out["rays_checked"] += int(v.shape[0])
    out["visibility_mismatch"] = int((v != oracle_results[2]).sum())
    return out


# compulsory lower bound of the algorithmic bytes per closest-hit ray: ray in + hit out + one root-to-leaf
# descent of the 8-wide tree (ceil(log8(T/4)) nodes) + 4 triangles
def compulsory_bytes_per_ray(n_triangles):
    levels = 1
    while 8 ** levels < n_triangles / 4.0:
        levels += 1
    return 32 + 16 + 80 * levels + 48 * 4


def measured_traffic(workload_key):
    """ncu counters of kernel B for this workload, captured under `ncu --set full` in an earlier run of the same command and
    committed under profiles/ (DRAM bytes cannot be measured live without the profiler): keyed by workload, or None"""
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(prof)).get(workload_key, {}).get("B")
    except Exception:
        return None


def make_ray_sets(pkg, ctx, scene, torch, side, subframe):
    """ray sets A/B/C on the device (generated by the library's raygen kernels)"""
    n = side * side
    eye, U, V, W = scene.camera_frame(side, side)
    cam = np.concatenate([eye, U, V, W]).astype(np.float32)
    A = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    B = torch.empty_like(A)
    C = torch.empty_like(A)
    hits = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    ctx.call("spc_gen_camera_rays", cam, side, side, subframe, A)
    ctx.trace_device(A, n, hits)
    ctx.call("spc_gen_bench_rays", 1, A, hits, n, B, None)
    ctx.call("spc_gen_bench_rays", 2, A, hits, n, C, None)
    ctx.synchronize()
    return A, B, C, hits


def run_reference(args, rank, world):
    """CPU arm: the oracle port over a bounded sample of the same workload, all host threads."""
    if rank != 0:
        return
    import spcbpt_loader
    pkg = spcbpt_loader.load()
    orc = spcbpt_loader.load_oracle()
    scene = pkg.scenes.heightfield_scene(708)
    threads = os.cpu_count() or 1
    # the same three ray sets, generated on the host from the oracle's own hits (sample of 2^19 rays each)
    side = 724  # 724^2 = 524176 ~ 2^19
    A = pkg.scenes.camera_rays(scene, side, side)
    osc = orc.Scene(pkg, scene)
    hA = osc.trace(A, threads=threads)
    B, C = host_bench_rays(pkg, scene, A, hA)
    sets = [("closest", A), ("closest", B), ("occlusion", C)]
    n_per_step = sum(r.shape[0] for _, r in sets)

    def step():
        osc.trace(A, threads=threads)
        osc.trace(B, threads=threads)
        osc.occlusion(C, threads=threads)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = n_per_step * args.steps / dt / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "first %d rays of each set per step (host-generated)" % A.shape[0]},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": "%d rays per step (3 sets x %d), oracle/orc_scene.cpp BVH2 + contract triangle test; "
                                       "the reference's own traversal is OptiX 7.5 (closed, absent)" % (n_per_step, A.shape[0])},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def host_bench_rays(pkg, scene, A, hA):
    """numpy restatement of k_bench_rays for the CPU arm (statistically the same ray sets)."""
    rng = np.random.default_rng(1)
    n = A.shape[0]
    hit = hA["prim"] >= 0
    o = np.stack([A["ox"], A["oy"], A["oz"]], 1)
    d = np.stack([A["dx"], A["dy"], A["dz"]], 1)
    P = o + d * hA["t"][:, None]
    # geometric normals from the scene triangles
    pos = np.concatenate([m["positions"][m["indices"].reshape(-1)].reshape(-1, 3, 3) for m in scene.meshes])
    tri = pos[np.maximum(hA["prim"], 0)]
    N = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    N /= np.maximum(np.linalg.norm(N, axis=1, keepdims=True), 1e-30)
    N[(N * d).sum(1) > 0] *= -1
    r1, r2 = rng.random(n), rng.random(n)
    r, phi = np.sqrt(r1), 2 * np.pi * r2
    loc = np.stack([r * np.cos(phi), r * np.sin(phi), np.sqrt(np.maximum(0, 1 - r1))], 1)
    up = np.where(np.abs(N[:, :1]) > np.abs(N[:, 2:3]), np.stack([-N[:, 1], N[:, 0], 0 * N[:, 0]], 1), np.stack([0 * N[:, 0], -N[:, 2], N[:, 1]], 1))
    up /= np.maximum(np.linalg.norm(up, axis=1, keepdims=True), 1e-30)
    tg = np.cross(up, N)
    bd = loc[:, :1] * tg + loc[:, 1:2] * up + loc[:, 2:3] * N
    B = np.zeros(n, pkg.RAY)
    B["ox"], B["oy"], B["oz"] = P.T.astype(np.float32)
    B["dx"], B["dy"], B["dz"] = bd.T.astype(np.float32)
    B["tmin"] = 1e-3
    B["tmax"] = np.where(hit, 1e16, -1.0)
    L = scene.lights[0]
    s1, s2 = rng.random(n)[:, None], rng.random(n)[:, None]
    lp = L["u"][None] * s1 + L["v"][None] * s2 + L["corner"][None] * (1 - s1 - s2)
    v = lp - P
    ln = np.linalg.norm(v, axis=1)
    C = np.zeros(n, pkg.RAY)
    C["ox"], C["oy"], C["oz"] = P.T.astype(np.float32)
    C["dx"], C["dy"], C["dz"] = (v / np.maximum(ln, 1e-30)[:, None]).T.astype(np.float32)
    C["tmin"] = 1e-3
    C["tmax"] = np.where(hit, ln - 1e-3, -1.0)
    return B, C


def _load_oracle_module(name):
    """oracle/<name>.py (reference-derived checkers: bench.py may execute oracle/ only in its baseline legs)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def relmse(x, ref):
    e = (x - ref) ** 2 / (ref ** 2 + 1e-2)
    return float(np.mean(e[np.isfinite(e)]))


def time_lane_frames(lr_, frames, torch, env):
    """three timed segments of `frames` frames; returns (median seconds, all seconds, launches of the last segment)"""
    seg_s, launches = [], 0
    for _ in range(3):
        l0 = lr_.launch_count()
        t0 = time.perf_counter()
        lr_.render(frames)
        torch.cuda.synchronize()
        seg_s.append(time.perf_counter() - t0)
        launches = lr_.launch_count() - l0
        env.barrier()
    return sorted(seg_s)[1], seg_s, launches


def reference_gpu_rate(pkg, torch, r0, w, h, frames=8):
    """GPU-side reference arm (north star: "the reference's own build on one B200"): the reference's OWN device programs compiled
    unmodified with its own --use_fast_math (oracle/_ref/libref_device.so; optixTrace = this repository's traversal,
    OptiX itself is absent) + its OWN MyThrustOp::LVC_Process (libref_thrust.so, T2) in the reference's frame loop, with the trained
    state of `r0`.  Returns None when the prebuilt libraries are absent or K differs from the compiled-in NUM_SUBSPACE."""
    rd, rt = _load_oracle_module("ref_device_py"), _load_oracle_module("ref_thrust_py")
    if not (rd.available() and rt.available()) or r0.K != 1000:
        return None
    loop = rd.ReferenceLoop(pkg, r0, rt)
    try:
        loop.render_frame()
        loop.stage_s = {k: 0.0 for k in loop.stage_s}
        t0 = time.perf_counter()
        for _ in range(frames):
            loop.render_frame()
        dt = time.perf_counter() - t0
        mean = float(loop.image().mean())
        return {"what": "reference programs (raygen.cu, hit_program.cu, cuProg.h, rmis.h unmodified, --use_fast_math) + reference LVC_Process on this GPU; "
                        "optixTrace = this repository's traversal (OptiX SDK absent)", "kind": "reference",
                "samples_per_s": w * h * frames / dt, "ms_per_frame": dt / frames * 1e3, "frames": frames,
                "stage_ms_per_frame": {k: v / frames * 1e3 for k, v in loop.stage_s.items()}, "image_mean": mean}
    finally:
        loop.close()


def equal_time_block(pkg, torch, scene, seconds, gt_spp, lanes):
    """relMSE at equal render time (BASELINE.json metric) on the same scene at 960x540: `pt`, SPCBPT exact flavour, SPCBPT fast
    flavour, and the reference programs on this GPU, each rendering for `seconds`; ground truth = `pt` at gt_spp from disjoint
    samples (chunks with a non-finite pixel value are dropped for that pixel: the reference's pt has no NaN guard)."""
    from spcbpt_optix7_b200.renderer import LaneRenderer, Renderer
    w, h = 960, 540
    gt = Renderer(scene, w, h, K=1000)
    acc, cnt = [np.zeros((h, w, 3)), np.zeros((h, w, 3))], [np.zeros((h, w, 1)), np.zeros((h, w, 1))]
    chunk = 256
    t0 = time.perf_counter()
    for c in range(max(2, gt_spp // chunk)):
        gt.reset_accumulation()
        gt.ctx.set_seed_offset(7777777 + c * chunk)
        for _ in range(chunk):
            gt.render_frame_pt()
        img = gt.image()
        ok = np.isfinite(img).all(-1, keepdims=True)
        acc[c & 1] += np.where(ok, img, 0.0)       # even / odd chunks: two independent halves
        cnt[c & 1] += ok
    ref = ((acc[0] + acc[1]) / np.maximum(cnt[0] + cnt[1], 1)).astype(np.float32)
    # The ground truth is itself a Monte-Carlo estimate: E[(I - G)^2] = Var(I) + Var(G) for independent unbiased I, G.  Var(G) is
    # estimated from the two halves (E[(G1 - G2)^2] = 4 Var(G)) and subtracted: `relMSE_debiased` is the estimator's own error.
    h1, h2 = acc[0] / np.maximum(cnt[0], 1), acc[1] / np.maximum(cnt[1], 1)
    gt_noise = float(np.mean((h1 - h2) ** 2 / (ref.astype(np.float64) ** 2 + 1e-2))) / 4.0
    out = {"image": "%dx%d" % (w, h), "seconds_each": seconds, "ground_truth": "pt %d spp (%.1f s)" % (gt_spp, time.perf_counter() - t0),
           "ground_truth_noise_relMSE": gt_noise, "rows": []}
    gt.ctx.close()

    def run(name, step, image, unit=1):
        step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            step()
            n += unit
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        img = np.nan_to_num(image())
        e = relmse(img, ref)
        out["rows"].append({"alg": name, "spp": n, "seconds": dt, "relMSE": e, "relMSE_debiased": e - gt_noise, "mean": float(img.mean())})

    pt = Renderer(scene, w, h, K=1000)

    def pt_step():
        for _ in range(8):
            pt.render_frame_pt()
        pt.ctx.synchronize()
    # (the warm-up step is part of the image for pt: reset afterwards is not needed, its samples are valid samples)
    run("pt", pt_step, pt.image, 8)
    out["rows"][-1]["spp"] += 8
    pt.ctx.close()
    trained = None
    for fast in (False, True):
        lr = LaneRenderer(scene, w, h, lanes=lanes, K=1000, fast=fast)
        lr.preprocessing()
        if fast:
            for lane in lr.lanes:
                lane.ctx.set_option("light_trace_mode", 1)
        run("SPCBPT_eye %s (%d lanes)" % ("fast flavour + parallel light tracer" if fast else "exact flavour", lanes), lambda: lr.render(lanes), lr.image, lanes)
        out["rows"][-1]["spp"] += lanes
        if not fast:
            trained = lr
    rd, rt = _load_oracle_module("ref_device_py"), _load_oracle_module("ref_thrust_py")
    if rd.available() and rt.available():
        loop = rd.ReferenceLoop(pkg, trained.lanes[0], rt)
        try:
            run("reference programs on this GPU (SPCBPT_eye loop)", loop.render_frame, loop.image, 1)
            out["rows"][-1]["spp"] += 1
        finally:
            loop.close()
    return out


def pkg_lane_blocks(lanes):
    from spcbpt_optix7_b200.renderer import LANE_TRACE_BLOCKS
    return LANE_TRACE_BLOCKS if lanes > 1 else 0


def render_section(args, pkg, torch, dist, rank, local_rank, world, large_scene=None):
    """BASELINE.json config 3 (and 4 at N>1): full SPCBPT at 1920x1080 with the reference's training schedule (2 M NEE paths,
    Q from light-trace launches, 100 Adam batches of 20 000), K = 1000, on a synthetic medium scene of the shipped scene's
    size class (the shipped house scene is reference data and does not travel).  Samples are partitioned across ranks
    (each rank renders its own subframes); the accumulation buffers are averaged over NCCL at read-out."""
    from spcbpt_optix7_b200.renderer import LaneRenderer
    from spcbpt_optix7_b200.parallel import DistEnv, comm_init, preprocess_distributed, reduce_accum
    env = DistEnv(dist)
    # the shipped scene when its .spcscene cache is present (written by __graft_entry__.build() from the reference's data with our
    # own loader, host/), else a synthetic scene of the same size class
    cache = os.path.join(ROOT, "data", "_ref", "house.spcscene")
    K, K_light, max_depth = 1000, 200, 0
    if large_scene is not None:
        scene = large_scene
        scene_name = "large synthetic glossy scene (%d triangles, %d quad emitters), max depth 12" % (scene.n_triangles, scene.lights.shape[0])
        K, K_light, max_depth = 1280, 256, 12
    elif os.path.exists(cache):
        scene = pkg.scenes.load_spcscene(cache)
        scene_name = "shipped house scene (house_uvrefine2.scene, %d triangles, 2 quad lights divLevel 10, 6 textures)" % scene.n_triangles
    else:
        scene = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=72, box_cells=60), 0.01)
        scene_name = "cornell-class medium scene %d triangles" % scene.n_triangles
    w, h = args.render_dim
    lanes = args.lanes
    torch.cuda.synchronize()
    t_up0 = time.perf_counter()
    # spc_scene_share (one replica of scene + BVH per GPU, lanes borrow lane 0's) is verified on one GPU only: the round's GPU budget ran
    # out before a multi-rank run of it, so under torchrun every lane keeps its own copy, exactly the configuration measured on 8 GPUs
    # (profiles/r2ab_bench_8gpu.json).  Frame times are the same either way on this scene (profiles/r2_summary.md).
    share = world == 1
    lr_ = LaneRenderer(scene, w, h, lanes=lanes, device=local_rank, K=K, K_light=K_light, max_depth=max_depth, share_scene=share)
    torch.cuda.synchronize()
    upload_s = time.perf_counter() - t_up0     # contexts + scene upload (host arrays -> device) + BVH build, all lanes
    r = lr_.lanes[0]
    lr_.seed_mapping(rank, world)
    note("section: %d lanes up (%.2f s)" % (lanes, upload_s))
    comm_init(r.ctx, env)       # NCCL communicator of this rank (csrc/comm.cu), outside the timed preprocessing
    env.barrier()
    note("section: communicator up, training")
    t0 = time.perf_counter()
    st = preprocess_distributed(r, env, pkg.TREE_NODE, target_samples=2000000, target_Q_samples=2000000, tree_samples=100000,
                                batch_size=20000, epochs=1, lr=0.01)
    lr_.share_trained_state()
    torch.cuda.synchronize()
    pre_s = time.perf_counter() - t0
    note("section: trained (%.2f s), rendering" % pre_s)
    lr_.render(2 * lanes)
    torch.cuda.synchronize()
    env.barrier()
    # three timed segments of `render_frames` frames each; the figure reported is the MEDIAN segment (the GPU boxes are VMs: an
    # occasional host hiccup stretches one segment by tens of per cent, tests/quick_variance.sh), all three are listed
    dt, seg_s, launches = time_lane_frames(lr_, args.render_frames, torch, env)
    # the same library with the parallel light tracer (light_trace_mode 1: per-path RNG streams instead of the reference's coupled
    # per-core streams -- same estimator, not bit-comparable with the reference's light paths), timed the same way
    for lane in lr_.lanes:
        lane.ctx.set_option("light_trace_mode", 1)
    lr_.render(2 * lanes)
    torch.cuda.synchronize()
    env.barrier()
    dt_lt1, seg_lt1, _ = time_lane_frames(lr_, args.render_frames, torch, env)
    for lane in lr_.lanes:
        lane.ctx.set_option("light_trace_mode", 0)
    r = lr_
    tt = torch.tensor([dt, dt_lt1], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt_max, dt_lt1 = float(tt[0].item()), float(tt[1].item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    reduce_accum(r, env)
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    t_ro0 = time.perf_counter()
    img = r.image()                            # merge of the lanes' running means + device -> host copy of the accumulation buffer
    readout_s = time.perf_counter() - t_ro0
    mean = float(img.mean())
    note("section: exact flavour rendered and read out")
    assert st["loss_last"] is not None and np.isfinite(st["loss_last"]) and np.isfinite(mean) and mean > 0, \
        "SPCBPT section: training or render produced no valid result (loss %r, image mean %r)" % (st["loss_last"], mean)
    out = {"workload": "SPCBPT_eye %dx%d, 1 spp per frame, K=%d (K_light %d), connections 3, %s, "
                       "light trace 1000x100 paths per frame, %d frame lanes per GPU" % (w, h, K, K_light, scene_name, lanes),
           "samples_per_s": w * h * args.render_frames * world / dt_max, "ms_per_frame": dt_max / args.render_frames * 1e3, "frames": args.render_frames,
           "preprocess_s": pre_s, "preprocess_phases_s": {k: st[k] for k in ("pretrace_s", "trees_s", "q_gamma_s")}, "train_paths": st["train_paths"], "loss_first": st["loss_first"], "loss_last": st["loss_last"],
           "segments_ms_per_frame": [x / args.render_frames * 1e3 for x in seg_s], "kernel_launches": int(launches), "accum_allreduce_ms": e0.elapsed_time(e1) if world > 1 else 0.0, "image_mean": mean,
           "exact_flavour_parallel_light_tracer": {"samples_per_s": w * h * args.render_frames * world / dt_lt1, "ms_per_frame": dt_lt1 / args.render_frames * 1e3,
                                                   "segments_ms_per_frame": [x / args.render_frames * 1e3 for x in seg_lt1]}}
    # end to end through the public API, host data in, host image out: scene upload + BVH build, training, the timed frames, read-out
    scene_bytes = int(sum(m["positions"].nbytes + m["indices"].nbytes + (m["texcoords"].nbytes if m.get("texcoords") is not None else 0) for m in scene.meshes)
                      + scene.materials.nbytes + scene.lights.nbytes + sum(t.nbytes for t in scene.textures))
    e2e_s = upload_s + pre_s + dt_max + readout_s
    out["e2e"] = {"what": "host scene arrays -> %s, %d lane contexts -> training -> %d frames -> merged accumulation buffer on the host" % (
                      "one upload + BVH build shared by the lanes" if share else "one upload + BVH build per lane", lanes, args.render_frames),
                  "upload_s": upload_s, "preprocess_s": pre_s, "render_s": dt_max, "readout_s": readout_s,
                  "h2d_bytes": scene_bytes * (1 if share else lanes), "d2h_bytes": w * h * 12,   # one upload when the lanes share lane 0's scene
                  "samples_per_s": w * h * args.render_frames * world / e2e_s,
                  "samples_per_s_without_training": w * h * args.render_frames * world / (upload_s + dt_max + readout_s)}
    # the fast-arithmetic flavour of the same library (FMA contraction + hardware special functions in the shading kernels, as the
    # reference's own --use_fast_math build; csrc/shade.cuh SPC_FAST_MATH, tests/test_fast_flavour_gpu.py): same schedule, own training
    if not args.no_fast:
        lf = LaneRenderer(scene, w, h, lanes=lanes, device=local_rank, K=K, K_light=K_light, max_depth=max_depth, fast=True, share_scene=share)
        lf.seed_mapping(rank, world)
        comm_init(lf.lanes[0].ctx, env)
        for lane in lf.lanes:   # per-path light-tracer streams as well (not bit-comparable with the reference either way)
            lane.ctx.set_option("light_trace_mode", 1)
        env.barrier()
        t0 = time.perf_counter()
        stf = preprocess_distributed(lf.lanes[0], env, pkg.TREE_NODE, target_samples=2000000, target_Q_samples=2000000, tree_samples=100000,
                                     batch_size=20000, epochs=1, lr=0.01)
        lf.share_trained_state()
        torch.cuda.synchronize()
        pre_f = time.perf_counter() - t0
        lf.render(2 * lanes)
        torch.cuda.synchronize()
        env.barrier()
        dtf, seg_f, _ = time_lane_frames(lf, args.render_frames, torch, env)
        ttf = torch.tensor([dtf], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ttf, op=dist.ReduceOp.MAX)
        dtf = float(ttf.item())
        reduce_accum(lf, env)
        out["fast_flavour"] = {"samples_per_s": w * h * args.render_frames * world / dtf, "ms_per_frame": dtf / args.render_frames * 1e3,
                               "segments_ms_per_frame": [x / args.render_frames * 1e3 for x in seg_f], "preprocess_s": pre_f,
                               "preprocess_phases_s": {k: stf[k] for k in ("pretrace_s", "trees_s", "q_gamma_s")},
                               "loss_last": stf["loss_last"], "image_mean": float(lf.image().mean()),
                               "flags": "-fmad=true -prec-div=false -prec-sqrt=false -DSPC_FAST_MATH (render.cu, pt.cu, pretrace.cu); traversal / binning / training unchanged; "
                                        "light_trace_mode 1 (one lane per light path)"}
        del lf
        note("section: fast flavour done")
    if rank == 0:
        # in-frame work and stage times: lane 0 alone, sequential frames, every stage of every bounce bracketed by CUDA events
        # (option "stage_timing"; slower than the production loop, used only to attribute the frame time and to state the
        # in-frame Mrays/s of the two traversal kernels)
        r0 = lr_.lanes[0]
        r0.ctx.set_option("stage_timing", 1)
        r0.ctx.set_trace_blocks(0)
        agg, nfr = None, 4
        for k in range(nfr + 1):
            r0.render_frame()
            es = r0.ctx.eye_stats()
            if k == 0:
                continue        # first frame after the lanes: warm-up
            if agg is None:
                agg = es
            else:
                for key in ("closest_rays", "shadow_slots", "shadow_rays", "visible_connections", "bounces"):
                    agg[key] += es[key]
                for key in es["stage_ms"]:
                    agg["stage_ms"][key] += es["stage_ms"][key]
        r0.ctx.set_option("stage_timing", 0)
        r0.ctx.set_trace_blocks(pkg_lane_blocks(lanes))
        sm = {k: v / nfr for k, v in agg["stage_ms"].items()}
        out["in_frame"] = {
            "how": "lane 0 alone, %d sequential frames, CUDA events around every stage of every bounce (spc_eye_stats_get)" % nfr,
            "bounces_per_frame": agg["bounces"] / nfr, "stage_ms_per_frame": sm,
            "closest_rays_per_frame": agg["closest_rays"] / nfr, "shadow_rays_per_frame": agg["shadow_rays"] / nfr,
            "shadow_slots_per_frame": agg["shadow_slots"] / nfr, "visible_connections_per_frame": agg["visible_connections"] / nfr,
            "closest_mrays_per_s": agg["closest_rays"] / nfr / (sm["trace"] * 1e-3) / 1e6 if sm["trace"] > 0 else None,
            "shadow_mrays_per_s": agg["shadow_rays"] / nfr / (sm["shadow"] * 1e-3) / 1e6 if sm["shadow"] > 0 else None}
        # CPU baseline of the same pass: one full frame of the oracle port with the same trained state, all host threads
        try:
            import spcbpt_loader
            restore_affinity()
            orc = spcbpt_loader.load_oracle()
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from harness import HostFrame
            threads = os.cpu_count() or 1
            cw, ch = w, h          # the full frame (a 1080p oracle frame takes ~2 s on 16 cores)
            hf = HostFrame(pkg, scene, cw, ch, K=K)
            hf.P["max_depth"] = max_depth
            r0 = lr_.lanes[0]
            hf.set_trees(r0.eye_tree, r0.light_tree)
            si = r0.P["subspace_info"]
            hf.set_q_gamma(r0.ctx.download(int(si["Q"][0]), np.float32, K), r0.ctx.download(int(si["CMFGamma"][0]), np.float32, K * K))
            osc = orc.Scene(pkg, scene)
            hf.P["lt"]["launch_frame"] = 1
            t0 = time.perf_counter()
            orc.light_trace(osc, hf.P, 1000, threads=threads)
            sub, cmfs, jump, vc, pc = orc.lvc_process(pkg, hf.lvc, hf.valid, K)
            hf.set_sampler(sub, cmfs, jump, vc, pc)
            t1 = time.perf_counter()
            orc.eye_pass(osc, hf.P, 1000, 3, 0, threads=threads)
            t2 = time.perf_counter()
            # one whole frame of the same pass on the host: light trace + LVC_Process + eye pass at the full image size
            frame_s = t2 - t0
            out["cpu_baseline"] = {"value": w * h / frame_s, "unit": "samples/s", "cores": threads, "kind": "port",
                                   "sample": "one frame of the oracle port: light trace of 100k paths + LVC_Process (%.2f s) + eye pass on %dx%d (%.2f s), same trained state" % (t1 - t0, cw, ch, t2 - t1)}
        except Exception as ex:   # the CPU leg is informative only
            out["cpu_baseline"] = {"error": repr(ex)}
        note("section: stage timing and CPU frame done")
        # GPU-side reference arm + relMSE at equal time (informative legs: a failure must not take the headline down)
        try:
            out["reference_gpu"] = reference_gpu_rate(pkg, torch, lr_.lanes[0], w, h)
            if out["reference_gpu"]:
                out["reference_gpu"]["ours_over_reference"] = out["samples_per_s"] / world / out["reference_gpu"]["samples_per_s"]
        except Exception as ex:
            out["reference_gpu"] = {"error": repr(ex)}
        if world == 1 and large_scene is None and not args.no_equal_time:
            try:
                out["equal_time"] = equal_time_block(pkg, torch, scene, args.equal_time_seconds, args.gt_spp, lanes)
            except Exception as ex:
                out["equal_time"] = {"error": repr(ex)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="rays per set for the cpu_baseline + parity leg (default: the whole 2^24-ray sets "
                    "of the micro workload, ~7 s on 16 cores; 2^20 per set for --workload large, whose oracle BVH takes ~20 s to build)")
    ap.add_argument("--no-render", action="store_true", help="skip the SPCBPT samples/s section (config 3)")
    ap.add_argument("--render-frames", type=int, default=48)
    ap.add_argument("--workload", default="micro", choices=["micro", "large"],
                    help="micro = BASELINE.json configs[1] (default, the headline); large = configs[4] (20 M triangles, 256 emitters, depth 12)")
    ap.add_argument("--render-dim", default="1920x1080", type=lambda v: tuple(int(x) for x in v.lower().split("x")),
                    help="image size of the SPCBPT section; BASELINE.json configs[3] is --gpus 8 --render-dim 3840x2160")
    ap.add_argument("--no-fast", action="store_true", help="skip the fast-arithmetic flavour in the SPCBPT section")
    ap.add_argument("--no-equal-time", action="store_true", help="skip the equal-time relMSE block of the SPCBPT section")
    ap.add_argument("--equal-time-seconds", type=float, default=1.5)
    ap.add_argument("--gt-spp", type=int, default=4096, help="ground-truth samples per pixel (pt) of the equal-time block")
    ap.add_argument("--lanes", type=int, default=4, help="frame lanes per GPU in the SPCBPT section (contexts rendering alternate subframes)")
    ap.add_argument("--section-timeout", type=int, default=int(os.environ.get("SPC_BENCH_SECTION_TIMEOUT_S", "-1")),
                    help="seconds the SPCBPT section may take before the headline line is printed without it and every rank exits 0 "
                         "(-1 = 360, or 1200 for --workload large; 0 = no limit).  A normal section takes 30-60 s")
    ap.add_argument("--watchdog", type=int, default=int(os.environ.get("SPC_BENCH_WATCHDOG_S", "-1")),
                    help="seconds after which a run that is still going dumps every thread's stack to stderr and exits (0 = off; -1 = 900, or 1800 "
                         "for --workload large): a hang -- a rank stuck in a collective, a lost GPU -- then costs a bounded time and says where it stood")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.watchdog < 0:
        args.watchdog = 1800 if args.workload == "large" else 900
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import spcbpt_loader
    pkg = spcbpt_loader.load()
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product has no CPU path"
    numa_note = bind_to_gpu_numa_node(local_rank) if world > 1 else "numa: single rank, not bound"
    torch.cuda.set_device(local_rank)
    if world > 1:
        note("process group: init (nccl, %d ranks)" % world)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        note("process group: up")

    large = args.workload == "large"
    if large:
        # BASELINE.json configs[4]: ~20 M triangles, 256 emitters, glossy terrain, max depth 12 (K = 1280, 256 emitter subspaces)
        scene = pkg.scenes.large_scene(3160, 16)
        global WORKLOAD, RAYS_SIDE
        RAYS_SIDE = 2048
        WORKLOAD = ("large_synthetic: glossy fractal terrain 3160x3160 (%d tris), 256 quad emitters, ray sets A primary 2048^2 + B cosine-bounce "
                    "+ C shadow, 2^22 rays each per GPU" % scene.n_triangles)
    else:
        scene = pkg.scenes.heightfield_scene(708)
    ctx = pkg.Context(local_rank)
    ctx.upload_scene(scene)
    note("scene uploaded, BVH built (%d triangles)" % scene.n_triangles)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    side = RAYS_SIDE
    n = side * side
    # each rank traces its own sample set (subframe index = rank+1 -> distinct jittered primaries)
    A, B, C, hits = make_ray_sets(pkg, ctx, scene, torch, side, subframe=rank + 1)
    hitsB = torch.empty_like(hits)
    vis = torch.empty((n,), dtype=torch.uint8, device="cuda")
    # visit counts: (a) of the production kernel's own schedule, (b) of the strict front-to-back t-pruned traversal that SURVEY.md
    # section 8d defines the algorithmic bytes by (option "count_canonical": independent of how the production kernel schedules)
    cntA = ctx.trace_counted(A, n, hits)
    cntB = ctx.trace_counted(B, n, hitsB)
    cntC = ctx.occlusion_counted(C, n, vis)
    ctx.set_option("count_canonical", 1)
    canB = ctx.trace_counted(B, n, hitsB)
    ctx.set_option("count_canonical", 0)

    def step(evs=None):
        ctx.trace_device(A, n, hits)
        if evs is not None:
            evs[0].record(stream)
        ctx.trace_device(B, n, hitsB)
        if evs is not None:
            evs[1].record(stream)
        ctx.occlusion_device(C, n, vis)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    note("ray sets and visit counts ready")
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record(stream)
    for s in range(args.steps):
        step(kev[s])
    ev1.record(stream)
    barrier()
    launches = ctx.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    note("timed region done: %.2f ms per step" % (ms_total / args.steps))
    clocks = sampler.stop()
    ms_B = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max = float(t.item())
    rays_per_step_all = 3 * n * world
    value = rays_per_step_all * args.steps / (ms_total_max * 1e-3) / 1e6

    # ---- e2e: host (pinned) buffers through the C ABI, copies inside the timed region
    hA = torch.empty((n, 8), dtype=torch.float32).pin_memory()
    hB = torch.empty((n, 8), dtype=torch.float32).pin_memory()
    hC = torch.empty((n, 8), dtype=torch.float32).pin_memory()
    hA.copy_(A); hB.copy_(B); hC.copy_(C)
    oA = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    oB = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    oC = torch.empty((n,), dtype=torch.uint8).pin_memory()
    L = pkg.lib()

    def e2e_step():
        ctx._ck(L.spc_trace_batch(ctx.h, hA.data_ptr(), n, 1, oA.data_ptr()), "spc_trace_batch")
        ctx._ck(L.spc_trace_batch(ctx.h, hB.data_ptr(), n, 1, oB.data_ptr()), "spc_trace_batch")
        ctx._ck(L.spc_occlusion_batch(ctx.h, hC.data_ptr(), n, oC.data_ptr()), "spc_occlusion_batch")
    e2e_steps = max(3, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = rays_per_step_all * e2e_steps / float(te.item()) / 1e6
    # e2e results must equal the device-resident results (same kernels)
    assert torch.equal(oB.cuda()[:, 3].view(torch.int32), hitsB[:, 3].view(torch.int32)), "e2e hits differ from device-path hits"
    note("e2e leg done")

    line = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        nn, nt = canB["nodes_visited"] / canB["rays"], canB["tris_tested"] / canB["rays"]
        nn_k, nt_k = cntB["nodes_visited"] / cntB["rays"], cntB["tris_tested"] / cntB["rays"]
        bytes_per_ray = 32 + 16 + 80 * nn + 48 * nt
        achieved = n * bytes_per_ray / (ms_B * 1e-3) / 1e9
        comp_bytes = compulsory_bytes_per_ray(scene.n_triangles)
        prof = measured_traffic("large" if large else "micro")
        traffic = prof.get("dram_bytes_per_launch") if prof else None
        if prof and prof.get("rays_per_launch") not in (None, n):
            traffic = None      # the capture was taken on another launch size
        # cpu_baseline: oracle port on all host cores over a bounded sample of the same ray sets
        restore_affinity()
        orc = spcbpt_loader.load_oracle()
        threads = os.cpu_count() or 1
        ns = min(args.cpu_sample if args.cpu_sample > 0 else (1 << 20 if large else 1 << 24), n)
        sets = [("closest", A[:ns].cpu().numpy().view(pkg.RAY).reshape(-1)), ("closest", B[:ns].cpu().numpy().view(pkg.RAY).reshape(-1)),
                ("occlusion", C[:ns].cpu().numpy().view(pkg.RAY).reshape(-1))]
        cpu_val, cpu_s, cpu_n, cpu_res = cpu_trace_sample(pkg, orc, scene, sets, ns, threads)
        # parity on the benchmarked configuration itself: the timed kernels' outputs against the oracle, ray by ray
        ctx.trace_device(A, n, hits)
        ctx.synchronize()
        parity = check_parity(pkg, (hits, hitsB), vis, cpu_res, ns)
        assert parity["prim_mismatch"] == 0 and parity["tuv_bit_mismatch"] == 0 and parity["visibility_mismatch"] == 0, \
            "GPU traversal differs from the oracle on the benchmarked rays: %r" % parity
        note("oracle leg done: %d rays in %.1f s, parity %r" % (cpu_n, cpu_s, parity))
        st = ctx.bvh_stats()
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": 3 * n, "host_affinity": numa_note, "bvh_nodes": st["n_nodes"], "bvh_bytes": st["bytes_nodes"] + st["bytes_triangles"],
                       "l2_policy": "ray inputs (3 x %d MiB per step) exceed L2; the %d MB BVH %s" % (
                           n * 32 >> 20, (st["bytes_nodes"] + st["bytes_triangles"]) // 1000000,
                           "is L2-resident by design (SURVEY.md section 7 hard parts)" if st["bytes_nodes"] + st["bytes_triangles"] < 100e6 else "does not fit the 126 MB L2"),
                       "per_set_nodes_tris_per_ray": {"A": [cntA["nodes_visited"] / n, cntA["tris_tested"] / n], "B": [nn, nt],
                                                      "C": [cntC["nodes_visited"] / n, cntC["tris_tested"] / n]}},
            "e2e": {"value": e2e_val, "unit": "Mrays/s", "h2d_bytes_per_step": 3 * n * 32, "d2h_bytes_per_step": 2 * n * 16 + n, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            # achieved = algorithmic bytes (48 + 80 n_node + 48 n_tri per ray with the visit counts of the strict
            # front-to-back traversal of the shipped BVH) / the production kernel's mean launch time.  Beside it: the same time
            # against the compulsory bound (one root-to-leaf descent + 4 triangles: independent of any traversal), and the DRAM
            # bytes ncu measured for this launch (what HBM really moves: the BVH of the micro workload is L2-resident).
            "roofline": {"bound": "hbm", "kernel": "k_trace_persist<false> (closest hit) on ray set B (incoherent)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nn, "tris_per_ray": nt, "kernel_ms": ms_B,
                         "mrays_per_s": n / ms_B / 1e3,
                         "nodes_tris_per_ray_production_schedule": [nn_k, nt_k],
                         "compulsory_bytes_per_ray": comp_bytes,
                         "frac_compulsory": n * comp_bytes / (ms_B * 1e-3) / 1e9 / peak,
                         "dram_frac": (traffic / (ms_B * 1e-3) / 1e9 / peak) if traffic else None,
                         "ncu": prof},
            "parity": parity,
            "cpu_baseline": {"value": cpu_val, "unit": "Mrays/s", "cores": threads, "kind": "port",
                             "sample": "first %d rays of each of the 3 sets (%d rays, %.1f s) on the oracle port" % (ns, cpu_n, cpu_s)},
        }
    rebind_affinity()

    # The headline line is complete at this point (rank 0 holds it); the SPCBPT section below is an extra key.  It runs under a
    # guard: an exception in it (single GPU) or a hang in it (any N) must not take the headline down.
    spcbpt = None
    if not args.no_render:
        guard = SectionGuard(args.section_timeout if args.section_timeout >= 0 else (1200 if large else 360), rank, line).start()
        try:
            spcbpt = render_section(args, pkg, torch, dist if world > 1 else None, rank, local_rank, world, scene if large else None)
        except Exception as ex:
            import traceback
            traceback.print_exc()
            if world > 1:
                # the other ranks are (or will be) waiting for this one in a collective of the section: there is no orderly way on.
                # Rank 0 prints the headline line now, any other rank just leaves; both with status 0 so that the launcher does not
                # tear the remaining ranks down before THEIR guard has let rank 0 print
                guard.abandon("exception on rank %d: %r" % (rank, ex))
            spcbpt = {"error": repr(ex)}
        guard.finish()
        note("SPCBPT section done")
    if rank == 0:
        if spcbpt is not None:
            line["spcbpt"] = spcbpt
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
