"""GPU parity against the REFERENCE'S OWN MyThrustOp library (cuda_thrust/device_thrust.cu compiled for sm_100a into
oracle/_ref/libref_thrust.so by oracle/Makefile; prebuilt in the authoring container, it travels to the GPU box).
Both libraries are fed the same device buffers (our pretrace / light-trace outputs) and every stage of the
post-processing seam is compared: LVC_Process, valid_sample_gather, sample_reweight, get_weighted_point_for_tree_building,
node_label, preprocess_getQ + Q_zero_handle, build_optimal_E_train_data, preprocess_getGamma (all bit-exact: the histogram
is summed per cell in the reference's serial order), train_optimal_E (thrust's own reductions have no specified order:
tolerance; ours is fixed, so a run is bit-reproducible) and Gamma2CMFGamma."""
import importlib.util
import os

import numpy as np
import pytest

from harness import DeviceFrame, compare_train, float_bits_differ, setup_pretrace_device

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rt():
    spec = importlib.util.spec_from_file_location("ref_thrust_py", os.path.join(ROOT, "oracle", "ref_thrust_py.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    if not m.available():
        pytest.skip("oracle/_ref/libref_thrust.so not prebuilt")
    return m


def test_post_processing_seam_vs_reference_library(gpu_ctx, rt):
    pkg = gpu_ctx
    K, KL = 1000, 200
    assert rt.lib().ref_thrust_num_subspace() == K
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=16, box_cells=10), 0.01)
    ctx = pkg.Context(0, K=K, K_light=KL)
    ctx.upload_scene(sc)
    cfg = dict(num_core=300, core_padding=400, M_per_core=50)
    df = DeviceFrame(pkg, sc, 640, 480, K=K, **cfg)
    n_core = 30000
    setup_pretrace_device(df, n_core, 10)
    # ---- valid_sample_gather over three pretrace launches
    for it in (1, 2, 3):
        df.P["pre_tracer"]["iteration"] = it
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_PRETRACE, n_core, 1)
        ctx.synchronize()
        a = ctx.valid_sample_gather(df.tp, n_core, df.tc, n_core * 10)
        b = rt.lib().ref_thrust_valid_sample_gather(df.tp.data_ptr(), n_core, df.tc.data_ptr(), n_core * 10)
        assert a == b > 10000
    gp, gc = ctx.train_set_read()
    rp, rc = rt.train_set_read(pkg)
    rp, rc = rp[:gp.shape[0]], rc[:gc.shape[0]]     # the reference over-allocates nothing, but be explicit
    assert not compare_train(pkg, gp, gc, rp, rc)
    assert np.array_equal(gc["path_id"], rc["path_id"])
    # ---- sample_reweight
    ctx.sample_reweight()
    rt.lib().ref_thrust_sample_reweight()
    gp, gc = ctx.train_set_read()
    rp, rc = rt.train_set_read(pkg)
    assert not float_bits_differ(gp["contri"], rp["contri"]).any()
    # ---- get_weighted_point_for_tree_building
    pts = {}
    for eye_side in (True, False):
        a = ctx.get_tree_points(eye_side, 20000)
        b = rt.tree_points(pkg, eye_side, 20000)
        assert a.shape == b.shape
        # light side: the reference leaves the record of an emitter endpoint uninitialised (device_thrust.cu:509-522)
        m = np.ones(a.shape[0], bool) if eye_side else (gc["light_source"][:a.shape[0]] == 0)
        for k in ("position", "dir", "normal", "weight"):
            assert not float_bits_differ(a[k][m], b[k][m]).any(), (eye_side, k)
        pts[eye_side] = a
    # ---- trees (host builder, bit-exact vs the reference's builder in the CPU tests) + node_label
    eye_tree, _ = pkg.build_tree(pts[True], K, 0)
    light_tree, _ = pkg.build_tree(pts[False], K - KL, 0)
    e_dev, l_dev = ctx.tree_to_device(True, eye_tree), ctx.tree_to_device(False, light_tree)
    re_dev = rt.lib().ref_thrust_tree_to_device(1, eye_tree.ctypes.data, eye_tree.shape[0])
    rl_dev = rt.lib().ref_thrust_tree_to_device(0, light_tree.ctypes.data, light_tree.shape[0])
    ctx.node_label(e_dev, l_dev)
    rt.lib().ref_thrust_node_label(re_dev, rl_dev)
    gp, gc = ctx.train_set_read()
    rp, rc = rt.train_set_read(pkg)
    assert np.array_equal(gc["label_A"], rc["label_A"]) and np.array_equal(gc["label_B"], rc["label_B"])
    # ---- LVC_Process + preprocess_getQ on two light-trace launches
    df.set_trees(eye_tree, light_tree)
    for frame in (1, 2):
        df.P["lt"]["launch_frame"] = frame
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_LIGHT_TRACE, cfg["num_core"], 1)
        ctx.synchronize()
        df.set_sampler_record(ctx.lvc_process(df.lvc, df.valid, df.n_lvc))
        gs, gcm, gj, gvc, gpc = df.sampler_host()
        rs, rcm, rj, rvc, rpc = rt.lvc_process(pkg, df.lvc.data_ptr(), df.valid.data_ptr(), df.n_lvc)
        assert (gvc, gpc) == (rvc, rpc) and gvc > 20000
        assert np.array_equal(gj, rj) and not float_bits_differ(gcm, rcm).any()
        for k in ("jump_bias", "id", "size"):
            assert np.array_equal(gs[k], rs[k]), k
        assert not float_bits_differ(gs["sum_pmf"], rs["sum_pmf"]).any()
        q_dev, acc = ctx.preprocess_getQ(df.lvc, df.valid, df.n_lvc, reset=(frame == 1))
        racc = rt.lib().ref_thrust_get_Q(df.lvc.data_ptr(), df.valid.data_ptr(), df.n_lvc, int(frame == 1))
        assert acc == racc
    ctx.Q_zero_handle()
    rt.lib().ref_thrust_Q_zero_handle()
    Qg, Qr = ctx.download(q_dev, np.float32, K), rt.download(rt.lib().ref_thrust_Q_ptr(), np.float32, K)
    assert not float_bits_differ(Qg, Qr).any()
    # ---- build_optimal_E_train_data (N a multiple of the batch size, as in the reference's own use)
    N = 40000
    assert gp.shape[0] >= N
    ctx.build_optimal_E_train_data(N)
    rt.lib().ref_thrust_build_train_data(N)
    gtd, rtd = ctx.train_data_read(), rt.train_data_read()
    assert gtd["N"] == rtd["N"] == N and gtd["M"] == rtd["M"]
    for k in ("P2N", "label_E", "label_P"):
        assert np.array_equal(gtd[k], rtd[k]), k
    for k in ("f_square", "pdf0", "peak"):
        assert not float_bits_differ(gtd[k], rtd[k]).any(), k
    # ---- preprocess_getGamma
    Gg = ctx.download(ctx.preprocess_getGamma(), np.float32, K * K).reshape(K, K)
    Gr = rt.download(rt.lib().ref_thrust_get_gamma(), np.float32, K * K).reshape(K, K)
    assert not float_bits_differ(Gg, Gr).any(), np.abs(Gg - Gr).max()   # same per-cell summation order as the reference's host loop
    Gg2 = ctx.download(ctx.preprocess_getGamma(), np.float32, K * K).reshape(K, K)
    assert np.array_equal(Gg.view(np.uint32), Gg2.view(np.uint32))
    # ---- train_optimal_E: 2 batches of 20000, lr .01 (the reference's constants)
    g_dev, loss = ctx.train_optimal_E(20000, 1, 0.01)
    Eg = ctx.download(g_dev, np.float32, K * K).reshape(K, K)
    Er = rt.download(rt.lib().ref_thrust_train_gamma(), np.float32, K * K).reshape(K, K)
    assert np.allclose(Eg.sum(1), 1, atol=1e-4) and np.allclose(Er.sum(1), 1, atol=1e-4)
    assert np.abs(Eg - Er).max() <= 2e-3 * Er.max(), (np.abs(Eg - Er).max(), Er.max())
    ctx.preprocess_getGamma()
    g_dev2, loss2 = ctx.train_optimal_E(20000, 1, 0.01)     # no floating-point atomics: a second run gives the same bits
    assert np.array_equal(ctx.download(g_dev2, np.float32, K * K).view(np.uint32), Eg.reshape(-1).view(np.uint32)) and np.array_equal(loss, loss2)
    # ---- Gamma2CMFGamma (each library on its own trained matrix)
    Cg = ctx.download(ctx.Gamma2CMFGamma(g_dev), np.float32, K * K).reshape(K, K)
    Cr = rt.download(rt.lib().ref_thrust_gamma_to_cmf(), np.float32, K * K).reshape(K, K)
    assert (Cg[:, -1] == 1).all() and (Cr[:, -1] == 1).all()
    assert np.abs(Cg - Cr).max() < 2e-3
