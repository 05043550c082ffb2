"""Multi-GPU host plumbing of the render core: one process per GPU, every rank holds a replica of scene + BVH (SURVEY.md
section 8e).  The data-path collectives are NCCL calls INSIDE libspcbpt_b200.so (csrc/comm.cu, include/spcbpt_b200.h
"multi-GPU"); this module only does what a host launcher does -- ship the NCCL unique id to the ranks and plan the shards:

  * training, once:  the NEE training paths and the Q light-trace launches are sharded across ranks (rank r traces pretrace
    iterations r+1, r+1+W, ... and light-trace frames likewise); the library all-reduces the reweighting grid, Q, the Gamma
    histogram and the K x K gradient of every Adam step, so every rank ends with the same Q / Gamma as a single-GPU run over
    the union of the shards (up to fp32 summation order); trees are built on rank 0's host and broadcast;
  * rendering:       rank r renders its own subframes (own light-vertex cache, own seeds) into its own running mean -- no
    per-bounce or per-frame collective;
  * read-out:        spc_reduce_accum sums the weighted running means into rank 0's buffer.

torch.distributed is used for the rendezvous only (broadcast of 128 id bytes, barriers, max of timings): with the gloo
backend on CPU this host logic is testable without GPUs (tests/test_parallel_cpu.py)."""
import numpy as np


class DistEnv:
    def __init__(self, dist=None):
        self.dist = dist if (dist is not None and dist.is_available() and dist.is_initialized()) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1

    def _dev(self):
        import torch
        return torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend() == "nccl" else torch.device("cpu")

    # -- rank-0 bytes to everyone (the NCCL unique id, tree arrays in the CPU tests) ----------------
    def broadcast_bytes(self, raw, nbytes=None):
        """raw: bytes on rank 0 (ignored elsewhere); returns rank 0's bytes on every rank"""
        if not self.dist or self.world == 1:
            return bytes(raw)
        import torch
        dev = self._dev()
        meta = torch.zeros(1, dtype=torch.int64, device=dev)
        if self.rank == 0:
            meta[0] = len(raw)
        self.dist.broadcast(meta, 0)
        n = int(meta[0].item())
        buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev) if self.rank == 0 else torch.empty(n, dtype=torch.uint8, device=dev)
        self.dist.broadcast(buf, 0)
        return buf.cpu().numpy().tobytes()

    def allreduce_mean(self, tensor):
        """in-place mean over ranks of a torch tensor (host-side statistics such as timings)"""
        if self.dist and self.world > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)
            tensor.div_(self.world)
        return tensor

    def barrier(self):
        if self.dist and self.world > 1:
            self.dist.barrier()


def shard_plan(rank, world, target_samples, target_Q_samples, batch_size):
    """who traces what: pure arithmetic, shared by renderer.py and host/spcbpt_main.cpp (tests/test_parallel_cpu.py)"""
    assert world >= 1 and 0 <= rank < world
    assert batch_size % world == 0, "the Adam batch (%d) must divide by the number of ranks (%d)" % (batch_size, world)
    return dict(rank=rank, world=world,
                local_samples=-(-target_samples // world),          # ceil: NEE training paths this rank traces
                local_Q_samples=-(-target_Q_samples // world),      # light paths behind this rank's Q estimate
                local_batch=batch_size // world,                    # this rank's share of every Adam batch
                first_iteration=rank + 1, iteration_stride=world,   # pretrace iterations rank+1, rank+1+W, ...
                first_lt_frame=rank + 1, lt_frame_stride=world,     # light-trace frames of the Q estimate likewise
                render_lt_base=1000003 * (rank + 1))                # light-trace frames of the render loop: disjoint from all of the above


def comm_init(ctx, env):
    """give `ctx` the NCCL communicator of this job: rank 0 creates the unique id, torch.distributed ships the 128 bytes"""
    if env.world == 1:
        return
    raw = ctx.comm_unique_id() if env.rank == 0 else b""
    ctx.comm_init(env.rank, env.world, env.broadcast_bytes(raw))


def preprocess_distributed(renderer, env, tree_dtype=None, **kw):
    """Renderer.preprocessing with the training set sharded over the ranks of `env` (collectives inside the library)"""
    if env.world > 1 and renderer.ctx.comm_info()[1] == 1:
        comm_init(renderer.ctx, env)
    plan = shard_plan(env.rank, env.world, kw.get("target_samples", 2000000), kw.get("target_Q_samples", 2000000), kw.get("batch_size", 20000))
    return renderer.preprocessing(plan=plan, **kw)


def reduce_accum(renderer, env, root=0):
    """read-out: every rank rendered the same number of subframes, so the image is the plain mean of the per-rank running means,
    summed into `root`'s buffer by NCCL (spc_reduce_accum).  `renderer` is a Renderer or a LaneRenderer (lanes merged first).
    Returns the accumulation tensor (valid on `root`; root < 0: on every rank)."""
    if hasattr(renderer, "merge"):
        renderer.merge()
    ctx = renderer.ctx
    ctx.synchronize()
    if env.world > 1:
        ctx.reduce_accum(renderer.accum, renderer.w * renderer.h, 1.0 / env.world, root)
    ctx.synchronize()
    return renderer.accum
