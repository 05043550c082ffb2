"""Multi-GPU plumbing of the render core: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch), every
rank holds a replica of the scene + BVH.  The path shards by samples: rank r renders its own
subframes (own light-vertex cache per frame, own seeds) into its own running mean; there is NO per-bounce or per-frame
collective.  Collectives appear in exactly two places:

  * training, once:  trees are built on rank 0's host and broadcast; Q [K], the Gamma histogram [K*K] and the trained
    matrix [K*K] are all-reduced (averaged) so that every rank samples from the same subspace statistics;
  * read-out:        the accumulation buffers [W*H*4] fp32 are all-reduced (averaged over ranks).

All helpers work with the gloo backend on CPU tensors as well, which is how the host logic is tested without GPUs."""
import numpy as np


class DistEnv:
    def __init__(self, dist=None):
        self.dist = dist if (dist is not None and dist.is_available() and dist.is_initialized()) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1

    # -- statistics: in-place average over ranks --------------------------------------------------
    def allreduce_mean(self, tensor):
        if self.dist and self.world > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM)
            tensor.div_(self.world)
        return tensor

    # -- rank-0 objects (numpy arrays) to everyone ------------------------------------------------
    def broadcast(self, obj):
        """broadcast(None) -> this rank; broadcast(obj) -> rank 0's obj (tuples of numpy arrays are sent as byte tensors)"""
        if obj is None:
            return self.rank
        if not self.dist or self.world == 1:
            return obj
        import torch
        backend_cuda = self.dist.get_backend() == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if backend_cuda else torch.device("cpu")
        out = []
        for k in range(len(obj)):
            a = obj[k]
            meta = torch.zeros(2, dtype=torch.int64, device=dev)
            if self.rank == 0:
                raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
                meta[0], meta[1] = raw.shape[0], a.dtype.itemsize
            self.dist.broadcast(meta, 0)
            n = int(meta[0].item())
            buf = torch.from_numpy(raw.copy()).to(dev) if self.rank == 0 else torch.empty(n, dtype=torch.uint8, device=dev)
            self.dist.broadcast(buf, 0)
            out.append(buf.cpu().numpy())
        return out

    def barrier(self):
        if self.dist and self.world > 1:
            self.dist.barrier()


def preprocess_distributed(renderer, env, tree_dtype, **kw):
    """Renderer.preprocessing with the two training exchanges wired to `env`: every rank traces its own training paths
    (pretrace iterations are offset by rank so that the sets are disjoint), rank 0 builds the trees."""
    renderer.P["pre_tracer"]["iteration"] = 1000003 * env.rank
    renderer.P["lt"]["launch_frame"] = 1000003 * env.rank

    def bcast(obj):
        r = env.broadcast(obj)
        if obj is None:
            return r
        return tuple(np.frombuffer(x.tobytes(), dtype=tree_dtype).copy() for x in r) if env.world > 1 else obj

    def allreduce(t):
        # the statistic was produced on the context's stream and is consumed there again; the collective runs on torch's
        # streams, so fence both sides (three times per training run: cost is irrelevant)
        renderer.ctx.synchronize()
        env.allreduce_mean(t)
        if getattr(t, "is_cuda", False):
            renderer.torch.cuda.synchronize(t.device)
    return renderer.preprocessing(allreduce=allreduce if env.world > 1 else None, broadcast=bcast if env.world > 1 else None, **kw)


def reduce_accum(renderer, env):
    """average the per-rank running means (every rank rendered the same number of subframes): the read-out collective.
    `renderer` is a Renderer or a LaneRenderer (whose lanes are merged first)."""
    if hasattr(renderer, "merge"):
        renderer.merge()
    renderer.ctx.synchronize()
    out = env.allreduce_mean(renderer.accum)
    if getattr(out, "is_cuda", False):
        renderer.torch.cuda.synchronize(out.device)
    return out
