"""Multi-GPU correctness on two B200s (skipped on a one-GPU box): one process per GPU, NCCL inside the library (csrc/comm.cu).
  * sharded training: two ranks each trace half of the NEE training paths and of the Q launches; the library all-reduces the
    reweighting grid, Q, the Gamma histogram and the K x K gradient of every Adam step.  The trained Gamma must equal -- up to fp32
    summation order -- a single-GPU training over the UNION of the two shards with the global batches made of both ranks' local
    batches (the union is rebuilt with spc_train_set_write / spc_train_Q_write);
  * sample-partitioned rendering + NCCL read-out: the reduced image of the two ranks equals the single-GPU render of the same
    global subframes with the same trained state (rtol 2e-6: the running mean is summed in another order)."""
import os
import sys
import tempfile
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, FRAMES = 96, 64, 6
KW = dict(K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)
TRAIN = dict(target_samples=80000, target_Q_samples=40000, tree_samples=20000, batch_size=20000)


def _scene(pkg):
    return pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)


def _worker(rank, world, id_file, out_dir):
    sys.path.insert(0, ROOT)
    import time
    import torch
    import spcbpt_loader
    pkg = spcbpt_loader.load()
    from spcbpt_optix7_b200.parallel import shard_plan
    from spcbpt_optix7_b200.renderer import Renderer
    torch.cuda.set_device(rank)
    r = Renderer(_scene(pkg), W, H, device=rank, **KW)
    if rank == 0:
        with open(id_file + ".tmp", "wb") as f:
            f.write(r.ctx.comm_unique_id())
        os.rename(id_file + ".tmp", id_file)
    else:
        for _ in range(600):
            if os.path.exists(id_file):
                break
            time.sleep(0.1)
    r.ctx.comm_init(rank, world, open(id_file, "rb").read())
    assert r.ctx.comm_info() == (rank, world)
    plan = shard_plan(rank, world, TRAIN["target_samples"], TRAIN["target_Q_samples"], TRAIN["batch_size"])
    st = r.preprocessing(plan=plan, **TRAIN)
    paths, conns = r.ctx.train_set_read()
    si = r.P["subspace_info"]
    Q = r.ctx.download(int(si["Q"][0]), np.float32, r.K)
    gamma = r.ctx.download(int(r.gamma_dev), np.float32, r.K * r.K)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), paths=paths, conns=conns, Q=Q, gamma=gamma, n_train=st["train_paths_used"],
             local_batch=plan["local_batch"], eye_tree=r.eye_tree, light_tree=r.light_tree, lt_base=plan["render_lt_base"])
    # sample-partitioned frames: this rank renders the global subframes rank, rank + world, ...
    r.ctx.set_seed_mapping(rank, world)
    for _ in range(FRAMES):
        r.render_frame()
    r.ctx.reduce_accum(r.accum, W * H, 1.0 / world, 0)
    if rank == 0:
        np.save(os.path.join(out_dir, "image.npy"), r.image())
        r.save_state(os.path.join(out_dir, "state_"))
    r.ctx.comm_barrier()


def test_two_ranks_sharded_training_and_reduced_image(gpu_ctx):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    pkg = gpu_ctx
    world = 2
    with tempfile.TemporaryDirectory() as d:
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_worker, args=(r, world, os.path.join(d, "nccl_id"), d)) for r in range(world)]
        for p in procs:
            p.start()
        deadline = time.monotonic() + 600
        for p in procs:
            p.join(timeout=max(1.0, deadline - time.monotonic()))
        stuck = [p for p in procs if p.is_alive()]
        for p in stuck:          # a rank waiting in a collective for a peer that failed: do not leave it (and its GPU) behind
            p.kill()
        for p in stuck:
            p.join(timeout=30)
        assert not stuck, "%d rank process(es) still running after 600 s" % len(stuck)
        for p in procs:
            assert p.exitcode == 0, "rank process failed"
        ranks = [np.load(os.path.join(d, "rank%d.npz" % r)) for r in range(world)]
        image2 = np.load(os.path.join(d, "image.npy"))
        # every rank ends with the same statistics
        assert np.array_equal(ranks[0]["Q"], ranks[1]["Q"]) and np.array_equal(ranks[0]["gamma"], ranks[1]["gamma"])
        assert np.array_equal(ranks[0]["eye_tree"], ranks[1]["eye_tree"]) and int(ranks[0]["n_train"]) == int(ranks[1]["n_train"])
        n_train, lb = int(ranks[0]["n_train"]), int(ranks[0]["local_batch"])
        assert n_train >= 2 * lb and lb == TRAIN["batch_size"] // world and lb >= 1000
        # ---- (a) single-GPU training over the union of the shards, global batch b = rank 0's batch b + rank 1's batch b -------
        chunks = []          # (rank, first path, last path) in global order
        for b in range(n_train // lb):
            for r in range(world):
                chunks.append((r, b * lb, (b + 1) * lb))
        for r in range(world):
            chunks.append((r, n_train, ranks[r]["paths"].shape[0]))     # paths beyond the trained range still feed the histogram
        paths, conns = [], []
        n_c = 0
        for r, lo, hi in chunks:
            p = ranks[r]["paths"][lo:hi].copy()
            if p.shape[0] == 0:
                continue
            c0, c1 = int(p["begin_ind"][0]), int(p["end_ind"][-1])
            c = ranks[r]["conns"][c0:c1].copy()
            shift = n_c - c0
            p["begin_ind"] += shift
            p["end_ind"] += shift
            c["path_id"] += sum(x.shape[0] for x in paths) - lo
            paths.append(p)
            conns.append(c)
            n_c += c.shape[0]
        paths, conns = np.concatenate(paths), np.concatenate(conns)
        assert (conns["path_id"][paths["begin_ind"]] == np.arange(paths.shape[0])).all()
        one = pkg.Context(0, K=KW["K"], K_light=KW["K_light"])
        one.upload_scene(_scene(pkg))
        one.train_set_write(paths, conns)
        one.train_Q_write(ranks[0]["Q"], 1)
        one.build_optimal_E_train_data(world * n_train)
        one.preprocess_getGamma()
        g_dev, loss = one.train_optimal_E(world * lb, 1, 0.01)
        gamma1 = one.download(g_dev, np.float32, KW["K"] ** 2)
        gamma2 = ranks[0]["gamma"]
        err = np.abs(gamma1 - gamma2).max() / np.abs(gamma1).max()
        print("sharded (2 ranks) vs single-GPU Gamma: max abs diff / max = %.3g, %d Adam steps" % (err, len(loss)))
        assert err < 1e-4, err
        one.close()
        # ---- (b) single-GPU render of the same global subframes with the same trained state --------------------------------------
        from spcbpt_optix7_b200.renderer import Renderer
        r = Renderer(_scene(pkg), W, H, device=0, **KW)
        r.load_state(os.path.join(d, "state_"))
        for g in range(world * FRAMES):
            rank, f = g % world, g // world
            r.P["lt"]["launch_frame"] = int(ranks[rank]["lt_base"]) + f      # launch_light_trace adds 1, as it did on that rank
            r.subframe = g
            r.render_frame()
        image1 = r.image()
        assert image1.mean() > 0.01
        assert np.allclose(image2, image1, rtol=2e-6, atol=1e-7), np.abs(image2 - image1).max()
