"""GPU parity of the subspace-training path (pretrace -> gather -> reweight -> trees -> labels -> Q -> training data ->
Gamma histogram -> Adam refinement -> CDF), through the C ABI, against the CPU oracle (oracle/orc_train.cpp, whose
pretrace part is pinned bit-for-bit to the reference's own __raygen__TrainData on the host shim).
Everything whose reference implementation has a defined summation order is compared bit-for-bit; the two stages
built on unordered reductions in the reference itself (thrust reduce_by_key / sort in the trainer, and our fp32
atomics in the Gamma histogram) are compared to the stated tolerances."""
import os

import numpy as np
import pytest

from harness import (DeviceFrame, HostFrame, compare_train, float_bits_differ, golden_scene, pretrace_host, setup_pretrace,
                     setup_pretrace_device)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_pretrace_vs_oracle_small(gpu_ctx, orc):
    pkg = gpu_ctx
    g = np.load(os.path.join(GOLD, "render.npz"))
    sc = golden_scene(pkg)
    K = 1000
    ctx = pkg.Context(0, K=K, K_light=200)
    ctx.upload_scene(sc)
    df = DeviceFrame(pkg, sc, 48, 40, K=K, num_core=16, core_padding=120, M_per_core=20)
    hf = HostFrame(pkg, sc, 48, 40, K=K, num_core=16, core_padding=120, M_per_core=20)
    for f in (df, hf):
        f.set_trees(g["eye_tree"], g["light_tree"])
    n = 6000
    setup_pretrace_device(df, n, 10, iteration=5)
    setup_pretrace(hf, n, 10, iteration=5)
    ctx.set_params(df.P)
    ctx.launch("pretrace", n, 1)
    ctx.synchronize()
    orc.pretrace(orc.Scene(pkg, sc), hf.P, K, threads=8)
    gp, gc = pretrace_host(df)
    bad = compare_train(pkg, gp, gc, hf.tp, hf.tc)
    assert not bad, bad
    assert gp["valid"].sum() > n // 3


@pytest.fixture(scope="module")
def chain(gpu_ctx, orc):
    """one training set built twice: on the GPU through the C ABI and by the oracle on the host"""
    pkg = gpu_ctx
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=16, box_cells=10), 0.01)   # ~5.6 units across, like the shipped house scene
    K, KL = 1000, 200
    ctx = pkg.Context(0, K=K, K_light=KL)
    ctx.upload_scene(sc)
    w, h = 640, 480
    cfg = dict(num_core=300, core_padding=400, M_per_core=50)
    df = DeviceFrame(pkg, sc, w, h, K=K, **cfg)
    hf = HostFrame(pkg, sc, w, h, K=K, **cfg)
    osc = orc.Scene(pkg, sc)
    ots = orc.TrainSet(pkg)
    n_core, n_iter = 30000, 3
    setup_pretrace_device(df, n_core, 10)
    setup_pretrace(hf, n_core, 10)
    counts = []
    for it in range(1, n_iter + 1):
        df.P["pre_tracer"]["iteration"] = it
        hf.P["pre_tracer"]["iteration"] = it
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_PRETRACE, n_core, 1)
        a = ctx.valid_sample_gather(df.tp, n_core, df.tc, n_core * 10)
        orc.pretrace(osc, hf.P, K, threads=8)
        b = ots.gather(hf.tp, hf.tc)
        counts.append((a, b))
    return dict(pkg=pkg, sc=sc, K=K, KL=KL, ctx=ctx, df=df, hf=hf, osc=osc, ots=ots, counts=counts, cfg=cfg)


def _cmp_sets(pkg, ga, gb, oa, ob):
    assert ga.shape == oa.shape and gb.shape == ob.shape
    bad = compare_train(pkg, ga, gb, oa, ob)
    assert np.array_equal(gb["path_id"], ob["path_id"])
    return bad


def test_gather_reweight_treepoints(chain, orc):
    c = chain
    pkg, ctx, ots = c["pkg"], c["ctx"], c["ots"]
    assert all(a == b for a, b in c["counts"]) and c["counts"][0][0] > 10000
    bad = _cmp_sets(pkg, *ctx.train_set_read(), *ots.read())
    assert not bad, bad
    ctx.sample_reweight()
    ots.reweight()
    gp, gc = ctx.train_set_read()
    op, oc = ots.read()
    d = float_bits_differ(gp["contri"], op["contri"])
    assert not d.any(), "reweighted contributions: %d differ" % d.sum()
    for eye_side in (True, False):
        a = ctx.get_tree_points(eye_side, 20000)
        b = ots.tree_points(eye_side, 20000)
        assert a.shape == b.shape and a.shape[0] > 20000
        for k in ("position", "dir", "normal", "weight"):
            assert not float_bits_differ(a[k], b[k]).any(), (eye_side, k)


def test_trees_labels_Q_traindata_gamma(chain, orc):
    c = chain
    pkg, ctx, ots, K, KL = c["pkg"], c["ctx"], c["ots"], c["K"], c["KL"]
    eye_tree, _ = pkg.build_tree(ctx.get_tree_points(True, 20000), K, 0)
    light_tree, _ = pkg.build_tree(ctx.get_tree_points(False, 20000), K - KL, 0)
    assert len(np.unique(eye_tree["label"])) > 500
    e_dev, l_dev = ctx.tree_to_device(True, eye_tree), ctx.tree_to_device(False, light_tree)
    df, hf = c["df"], c["hf"]
    for f in (df, hf):
        f.set_trees(eye_tree, light_tree)
    ctx.node_label(e_dev, l_dev)
    ots.label(eye_tree, light_tree)
    gp, gc = ctx.train_set_read()
    op, oc = ots.read()
    assert np.array_equal(gc["label_A"], oc["label_A"]) and np.array_equal(gc["label_B"], oc["label_B"])
    # Q from two light-trace launches (preprocess_getQ)
    oq = orc.QEstimator(K)
    for frame in (1, 2):
        df.P["lt"]["launch_frame"] = frame
        hf.P["lt"]["launch_frame"] = frame
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_LIGHT_TRACE, c["cfg"]["num_core"], 1)
        q_dev, acc = ctx.preprocess_getQ(df.lvc, df.valid, df.n_lvc, reset=(frame == 1))
        orc.light_trace(c["osc"], hf.P, K, threads=8)
        oacc = oq.add(hf.lvc, hf.valid)
        assert acc == oacc
    ctx.Q_zero_handle()
    oq.zero_handle()
    Qg, Qo = ctx.download(q_dev, np.float32, K), oq.read()
    assert not float_bits_differ(Qg, Qo).any(), "Q: %d of %d entries differ" % (float_bits_differ(Qg, Qo).sum(), K)
    assert (Qg < 1e30).sum() > 100
    # training arrays
    n = gp.shape[0]
    ctx.build_optimal_E_train_data(n)
    otd = ots.build_train_data(n, Qo, K)
    gtd = ctx.train_data_read()
    assert gtd["N"] == otd["N"] and gtd["M"] == otd["M"]
    assert np.float32(gtd["threshold"]).view(np.uint32) == np.float32(otd["threshold"]).view(np.uint32)
    for k in ("P2N", "label_E", "label_P"):
        assert np.array_equal(gtd[k], otd[k]), k
    for k in ("f_square", "pdf0", "peak"):
        assert not float_bits_differ(gtd[k], otd[k]).any(), k
    # Gamma histogram: our scatter-add uses fp32 atomics (unordered); tolerance 1e-5 relative on the normalised rows
    g_dev = ctx.preprocess_getGamma()
    Gg = ctx.download(g_dev, np.float32, K * K).reshape(K, K)
    Go = ots.gamma_histogram(K)
    assert np.allclose(Gg, Go, rtol=1e-5, atol=1e-9), np.abs(Gg - Go).max()
    assert np.allclose(Gg.sum(1), 1, atol=1e-4)
    # Adam refinement: 2 batches of 20000 (the reference's batch size); losses to 1e-4 relative, E to 2e-3 absolute of its
    # scale.  (Adam's first steps move every theta by ~lr regardless of gradient size, so tiny sum-order differences in the
    # gradient can flip steps of near-zero entries: the tolerance is on E, which those entries barely affect.)
    g_dev, loss_g = ctx.train_optimal_E(20000, 1, 0.01)
    Eo, loss_o = orc.train_gamma(otd, K, Go, 20000, 1, 0.01)
    Eg = ctx.download(g_dev, np.float32, K * K).reshape(K, K)
    assert loss_g.shape == loss_o.shape and loss_g.shape[0] == n // 20000 >= 2
    assert np.allclose(loss_g, loss_o, rtol=1e-4), (loss_g, loss_o)
    assert np.allclose(Eg.sum(1), 1, atol=1e-4)
    assert np.abs(Eg - Eo).max() <= 2e-3 * Eo.max(), (np.abs(Eg - Eo).max(), Eo.max())
    # CDF
    cmf_dev = ctx.Gamma2CMFGamma(g_dev)
    Cg = ctx.download(cmf_dev, np.float32, K * K).reshape(K, K)
    Co = orc.gamma_to_cmf(Eg, K)
    assert not float_bits_differ(Cg, Co).any()
    assert (Cg[:, -1] == 1).all() and (np.diff(Cg, axis=1) >= 0).all()
