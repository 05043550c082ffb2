"""CPU tests (gloo, world_size 2) of the multi-GPU host logic in spcbpt-optix7_b200/parallel.py: the rendezvous that ships the
NCCL unique id (rank-0 bytes reach every rank bit-for-bit), host-side statistics, and the shard plan (which rank traces which
pretrace iterations / light-trace frames, how the Adam batch splits).  The data-path collectives themselves are NCCL calls inside
the library and are tested on GPUs (tests/test_multigpu_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import spcbpt_loader
    pkg = spcbpt_loader.load()
    from spcbpt_optix7_b200.parallel import DistEnv
    env = DistEnv(dist)
    ok = env.rank == rank and env.world == world
    # statistics: Q-like vector and a Gamma-like matrix
    qv = torch.full((16,), float(rank + 1))
    env.allreduce_mean(qv)
    ok &= bool(torch.allclose(qv, torch.full((16,), (1 + world) / 2)))
    # trees: built on rank 0 only
    if rank == 0:
        g = np.random.default_rng(1)
        s = np.zeros(500, pkg.DIVIDE_WEIGHT)
        s["position"] = g.uniform(-1, 1, (500, 3))
        s["normal"] = (0, 1, 0)
        s["weight"] = g.uniform(0.1, 1, 500)
        eye, _ = pkg.build_tree(s, 8, 0)
        light, _ = pkg.build_tree(s, 6, 0)
        obj = (eye, light)
    else:
        obj = (np.zeros(0, pkg.TREE_NODE), np.zeros(0, pkg.TREE_NODE))
    got = [env.broadcast_bytes(x.tobytes() if rank == 0 else b"") for x in obj]
    trees = [np.frombuffer(x, dtype=pkg.TREE_NODE) for x in got]
    # the 128-byte NCCL unique id travels the same way (parallel.comm_init)
    uid = bytes(range(128)) if rank == 0 else b""
    ok &= env.broadcast_bytes(uid) == bytes(range(128))
    digest = [int(t["label"].sum()) + int(t["child"].sum()) + t.shape[0] for t in trees]
    # accumulation buffers: mean over ranks
    acc = torch.full((32, 4), float(rank))
    env.allreduce_mean(acc)
    ok &= bool(torch.allclose(acc, torch.full((32, 4), (world - 1) / 2)))
    env.barrier()
    q.put((rank, ok, digest))
    dist.destroy_process_group()


def test_distenv_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] for r in res)
    assert res[0][2] == res[1][2] and res[0][2][0] > 10     # identical trees on both ranks


def test_shard_plan_partitions_the_work():
    sys.path.insert(0, ROOT)
    import spcbpt_loader
    spcbpt_loader.load()
    from spcbpt_optix7_b200.parallel import shard_plan
    for world in (1, 2, 4, 8):
        plans = [shard_plan(r, world, 2000000, 2000000, 20000) for r in range(world)]
        assert sum(p["local_batch"] for p in plans) == 20000
        assert sum(p["local_samples"] for p in plans) >= 2000000 and plans[0]["local_samples"] == -(-2000000 // world)
        # pretrace iterations / Q light-trace frames of the ranks are disjoint and cover 1, 2, 3, ...
        its = sorted(p["first_iteration"] + k * p["iteration_stride"] for p in plans for k in range(5))
        assert its == list(range(1, 5 * world + 1))
        fr = sorted(p["first_lt_frame"] + k * p["lt_frame_stride"] for p in plans for k in range(5))
        assert fr == list(range(1, 5 * world + 1))
        # render-loop light-trace frames never collide with training frames or with another rank's
        bases = [p["render_lt_base"] for p in plans]
        assert len(set(bases)) == world and min(bases) > 100000
    one = shard_plan(0, 1, 2000000, 2000000, 20000)
    assert (one["local_samples"], one["local_batch"], one["first_iteration"], one["iteration_stride"]) == (2000000, 20000, 1, 1)
    try:
        shard_plan(0, 3, 2000000, 2000000, 20000)
        assert False, "20000 does not divide by 3"
    except AssertionError as ex:
        assert "divide" in str(ex)


def test_tile_partition_formula_is_the_reference_work_distribution():
    """The GPU test of spc_set_tile_partition (tests/test_pipeline_gpu.py) checks that rank g renders exactly the pixels with
    ((x // 8) - (y // 4)) mod G == g.  Here that closed form is compared with the reference's own StaticWorkDistribution
    (sutil/WorkDistribution.h:34-91, compiled from the reference tree): its getSamplePixel enumeration over all GPUs is a partition of
    the image with exactly that ownership, also when the image is no multiple of the tile strip."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("spc_ref_py", os.path.join(root, "oracle", "ref_py.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    if not (ref.available() and hasattr(ref.lib(), "ref_tile_owner_map")):
        pytest.skip("oracle/_ref/libref_host.so without ref_tile_owner_map (built where /root/reference exists)")
    for w, h in ((1920, 1080), (3840, 2160), (96, 64), (100, 37), (7, 5), (1, 1), (64, 3)):
        for G in (1, 2, 3, 4, 8):
            owner, twice = ref.tile_owner_map(w, h, G)
            yy, xx = np.mgrid[0:h, 0:w]
            assert twice == 0 and (owner >= 0).all(), (w, h, G)
            assert np.array_equal(owner, (xx // 8 - yy // 4) % G), (w, h, G)
