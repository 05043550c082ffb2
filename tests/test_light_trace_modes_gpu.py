"""The parallel light tracer (spc_set_option "light_trace_mode" 1: one lane per light path, per-path RNG streams, vertices packed
densely in path order by a count / scan / write double pass) against the reference-stream mode (one serial core per lane,
raygen.cu:620-685).  Different random numbers, same estimator: the two are compared statistically; mode 1 itself must be
bit-reproducible and must keep the LVC contract LVC_Process relies on (valid prefix, depth-0 vertex opens every path)."""
import numpy as np
import pytest

from harness import DeviceFrame, random_trees_and_gamma

pytestmark = pytest.mark.gpu


def test_parallel_light_tracer_contract_and_statistics(gpu_ctx):
    pkg = gpu_ctx
    sc = pkg.scenes.cornell_scene(wall_cells=12, box_cells=8)
    K, KL = 64, 12
    cfg = dict(num_core=400, core_padding=400, M_per_core=50)
    P = np.concatenate([m["positions"][m["indices"].astype(np.int64)].mean(1) for m in sc.meshes]).astype(np.float32)
    N = np.tile(np.array([[0, 1, 0]], np.float32), (P.shape[0], 1))
    eye_tree, light_tree, Q, cmf = random_trees_and_gamma(pkg, P, N, K, KL, lambda p, s, k, b: p.build_tree(s, k, b), seed=2)
    ctx = pkg.Context(0, K=K, K_light=KL)
    ctx.upload_scene(sc)
    df = DeviceFrame(pkg, sc, 32, 32, K=K, **cfg)
    df.P["subspace_info"]["eye_tree"] = ctx.tree_to_device(True, eye_tree)
    df.P["subspace_info"]["light_tree"] = ctx.tree_to_device(False, light_tree)
    df.set_q_gamma(Q, cmf)

    def run(mode, frame):
        ctx.set_option("light_trace_mode", mode)
        df.P["lt"]["launch_frame"] = frame
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_LIGHT_TRACE, cfg["num_core"], 1)
        ctx.synchronize()
        return df.lvc_host()

    n_paths = cfg["num_core"] * cfg["M_per_core"]
    a, va = run(1, 5)
    b, vb = run(1, 5)
    assert np.array_equal(va, vb) and np.array_equal(a[va == 1].view(np.uint8), b[vb == 1].view(np.uint8)), "mode 1 is not reproducible"
    c, vc = run(1, 6)
    assert not np.array_equal(a["position"][:1000], c["position"][:1000]), "the launch frame must change the paths"
    # contract: the valid slots are a dense prefix, every path opens with its depth-0 emitter vertex, depths count up along a path
    nv = int(va.sum())
    assert va[:nv].all() and not va[nv:].any()
    v = a[:nv]
    assert int((v["depth"] == 0).sum()) == n_paths and v["depth"][0] == 0
    step = np.diff(v["depth"].astype(np.int32))
    assert ((step == 1) | (v["depth"][1:] == 0)).all()
    assert (v["isOrigin"][v["depth"] == 0] == 1).all() and (v["subspaceId"][v["depth"] == 0] >= K - KL).all()
    assert (v["subspaceId"][v["depth"] > 0] < K - KL).all() and np.isfinite(v["flux"]).all()
    # statistics against the reference-stream mode (independent samples of the same distribution): vertices per path, the depth
    # histogram and the per-subspace flux/pdf mass (the Q vector of preprocess_getQ)
    s0, s1 = [], []
    for frame in range(10, 14):
        for mode, acc in ((0, s0), (1, s1)):
            x, vx = run(mode, frame)
            x = x[vx == 1]
            w = x["flux"].sum(1) / x["pdf"]
            w[~np.isfinite(w)] = 0
            acc.append((x.shape[0], np.bincount(np.minimum(x["depth"], 12), minlength=13), np.bincount(x["subspaceId"], weights=w, minlength=K)))
    n0, n1 = sum(t[0] for t in s0), sum(t[0] for t in s1)
    d0, d1 = sum(t[1] for t in s0), sum(t[1] for t in s1)
    q0, q1 = sum(t[2] for t in s0), sum(t[2] for t in s1)
    print("vertices: serial cores %d, parallel paths %d; depth histograms %s / %s" % (n0, n1, d0[:6], d1[:6]))
    assert abs(n1 / n0 - 1) < 0.02
    # two independent Poisson-like counts per depth: 4.5 standard deviations of their difference
    assert (np.abs(d1[:6].astype(float) - d0[:6]) <= 4.5 * np.sqrt(d0[:6] + d1[:6] + 1.0)).all()
    assert abs(q1.sum() / q0.sum() - 1) < 0.02
    big = q0 > 0.01 * q0.sum()
    assert big.sum() >= 5 and np.allclose(q1[big], q0[big], rtol=0.1)
    ctx.set_option("light_trace_mode", 0)
    ctx.close()


def test_parallel_light_tracer_renders_the_same_image(gpu_ctx):
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    kw = dict(K=64, K_light=12, lt_num_core=200, lt_core_padding=400, lt_M_per_core=50, pretrace_num_core=20000)
    r = Renderer(sc, 128, 96, **kw)
    r.preprocessing(target_samples=80000, target_Q_samples=60000, tree_samples=30000, batch_size=20000)
    imgs = []
    for mode in (0, 1):
        r.ctx.set_option("light_trace_mode", mode)
        r.reset_accumulation()
        for _ in range(256):
            r.render_frame()
        imgs.append(r.image().copy())
    a, b = imgs
    relmse = float(np.mean((a - b) ** 2 / (a ** 2 + 1e-2)))
    print("serial cores vs parallel paths at 256 spp: means %.5f / %.5f, relMSE between them %.5f" % (a.mean(), b.mean(), relmse))
    assert np.isfinite(b).all() and abs(a.mean() / b.mean() - 1) < 0.01 and relmse < 0.01
