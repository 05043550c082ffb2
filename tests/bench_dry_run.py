"""CPU dry run of bench.main() (test helper, run by tests/test_bench_contract.py): torch.cuda is replaced by a fake surface and
pkg.Context by a fake context that answers the traversal calls through the oracle, on a tiny scene and 48 x 48 rays per set.  It
exercises bench.py's control flow -- the order of the legs, the assembly and JSON encoding of the headline line, the parity
check, the guard around the SPCBPT section -- and nothing of the product (the numbers it prints mean nothing).
usage: python tests/bench_dry_run.py norender|ok|raise|hang
Under torchrun (WORLD_SIZE > 1; the process group is switched to gloo) the modes ok2 | hang1 | raise1 run the multi-rank flow: in
hang1 / raise1 rank 1 gets stuck / fails inside the section while rank 0 waits for it in a collective."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import spcbpt_loader
pkg = spcbpt_loader.load()
orc = spcbpt_loader.load_oracle()

mode = sys.argv[1]
bench.RAYS_SIDE = 48

# ---- fake torch.cuda ----
class FakeStream:
    cuda_stream = 0
class FakeEvent:
    def __init__(self, enable_timing=False): pass
    def record(self, s=None): self.t = time.perf_counter()
    def elapsed_time(self, other): return max((other.t - self.t) * 1e3, 1e-3)
torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda d: None
torch.cuda.current_stream = lambda *a: FakeStream()
torch.cuda.Event = FakeEvent
torch.cuda.synchronize = lambda *a: None
_empty, _tensor = torch.empty, torch.tensor
torch.empty = lambda *a, **k: _empty(*a, **{x: y for x, y in k.items() if x != "device"})
torch.tensor = lambda *a, **k: _tensor(*a, **{x: y for x, y in k.items() if x != "device"})
torch.Tensor.pin_memory = lambda self: self
torch.Tensor.cuda = lambda self: self

class FakeLib:
    def __init__(self, ctx): self.ctx = ctx
    def spc_trace_batch(self, h, rays_ptr, n, flags, out_ptr):
        self.ctx.by_ptr_trace(rays_ptr, n, out_ptr); return 0
    def spc_occlusion_batch(self, h, rays_ptr, n, out_ptr):
        self.ctx.by_ptr_occ(rays_ptr, n, out_ptr); return 0

class FakeContext:
    h = 0
    def __init__(self, dev): self.launches = 0; self.tensors = {}
    def upload_scene(self, scene): self.scene = scene; self.osc = orc.Scene(pkg, scene)
    def set_stream(self, s): pass
    def synchronize(self): pass
    def launch_count(self): return self.launches
    def set_option(self, n, v): pass
    def _ck(self, rc, what): assert rc == 0
    def bvh_stats(self): return {"n_nodes": 10, "bytes_nodes": 800, "bytes_triangles": 4800}
    def reg(self, t): self.tensors[t.data_ptr()] = t; return t
    def call(self, name, *args):
        if name == "spc_gen_camera_rays":
            cam, w, h, sub, A = args
            A.copy_(torch.from_numpy(pkg.scenes.camera_rays(self.scene, w, h).view(np.float32).reshape(-1, 8)))
            self.A = A
        elif name == "spc_gen_bench_rays":
            kind, A, hits, n, out, _ = args
            a = A.numpy().view(pkg.RAY).reshape(-1); hA = hits.numpy().view(pkg.HIT).reshape(-1)
            B, C = bench.host_bench_rays(pkg, self.scene, a, hA)
            out.copy_(torch.from_numpy((B if kind == 1 else C).view(np.float32).reshape(-1, 8)))
    def trace_device(self, rays, n, hits):
        self.reg(rays); self.reg(hits)
        r = self.osc.trace(rays.numpy().view(pkg.RAY).reshape(-1)[:n])
        hits[:n].copy_(torch.from_numpy(r.view(np.float32).reshape(-1, 4))); self.launches += 1
    def occlusion_device(self, rays, n, vis):
        self.reg(rays); self.reg(vis)
        vis[:n].copy_(torch.from_numpy(self.osc.occlusion(rays.numpy().view(pkg.RAY).reshape(-1)[:n]).astype(np.uint8))); self.launches += 1
    def trace_counted(self, rays, n, hits):
        self.trace_device(rays, n, hits); return {"nodes_visited": 17 * n, "tris_tested": 7 * n, "rays": n}
    def occlusion_counted(self, rays, n, vis):
        self.occlusion_device(rays, n, vis); return {"nodes_visited": 9 * n, "tris_tested": 3 * n, "rays": n}
    def by_ptr_trace(self, rp, n, op):
        self.trace_device(self.tensors[rp], n, self.tensors[op])
    def by_ptr_occ(self, rp, n, op):
        self.occlusion_device(self.tensors[rp], n, self.tensors[op])

the_ctx = {}
def make_ctx(dev):
    c = FakeContext(dev); the_ctx["c"] = c; return c
pkg.Context = make_ctx
pkg.lib = lambda *a: FakeLib(the_ctx["c"])
# e2e leg addresses tensors by data_ptr: register pinned tensors as they are created
_orig_empty = torch.empty
def reg_empty(*a, **k):
    t = _orig_empty(*a, **k)
    if "c" in the_ctx: the_ctx["c"].reg(t)
    return t
torch.empty = reg_empty
pkg.scenes.heightfield_scene = lambda n: pkg.scenes.cornell_scene(wall_cells=4, box_cells=3)

import torch.distributed as tdist
_init = tdist.init_process_group
tdist.init_process_group = lambda backend, **kw: _init("gloo")      # no NCCL without GPUs


def fake_section(args, pkg_, torch_, dist, rank, local_rank, world, large_scene=None):
    if mode == "hang": time.sleep(60)
    if mode == "raise": raise RuntimeError("boom in section")
    if mode == "hang1" and rank == 1: time.sleep(60)
    if mode == "raise1" and rank == 1: raise RuntimeError("boom on rank 1")
    if dist is not None:
        dist.barrier()          # the section's collectives: rank 0 waits here for a rank that never comes
    return {"samples_per_s": 123.0}
bench.render_section = fake_section
argv = ["bench.py", "--steps", "2", "--warmup", "3"]
if mode == "norender": argv.append("--no-render")
if mode in ("hang", "hang1", "raise1"): argv += ["--section-timeout", "3"]
sys.argv = argv
bench.main()
