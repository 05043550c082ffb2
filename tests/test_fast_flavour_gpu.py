"""The fast-arithmetic flavour (libspcbpt_b200_fast.so: FMA contraction, approximate division / square root, hardware fp32
special functions in the shading kernels -- what the reference's own --use_fast_math build does, src/CMakeLists.txt:214-215)
against the exact flavour, at the tolerance BASELINE.json's north star states:
  * closest-hit primitive ids and t/u/v: bit-equal (the traversal and the primary rays are the same code in both);
  * first-vertex subspace ids: equal (a vertex within an ulp of a split plane may flip: <= 1e-5 of the pixels);
  * shading, stage-wise on identical inputs (same LVC, same sampler, subframe 0, one bounce): per-pixel radiance within 1e-5
    relative for the bulk of the pixels (discrete decisions -- Russian roulette, the sampled light vertex -- flip for a few);
  * converged images: equal means and a relMSE between the two flavours far below the noise floor of either."""
import numpy as np
import pytest

from harness import DeviceFrame, random_trees_and_gamma, varied_cornell

pytestmark = pytest.mark.gpu


def _setup(pkg, fast, sc, K, KL, w, h, cfg, trees):
    ctx = pkg.Context(0, K=K, K_light=KL, connections=3, fast=fast)
    ctx.upload_scene(sc)
    df = DeviceFrame(pkg, sc, w, h, K=K, **cfg)
    eye_tree, light_tree, Q, cmf = trees
    df.P["subspace_info"]["eye_tree"] = ctx.tree_to_device(True, eye_tree)
    df.P["subspace_info"]["light_tree"] = ctx.tree_to_device(False, light_tree)
    df.set_q_gamma(Q, cmf)
    return ctx, df


def test_fast_flavour_stagewise_against_exact(gpu_ctx):
    import torch
    pkg = gpu_ctx
    sc = varied_cornell(pkg)
    K, KL, w, h = 1000, 200, 192, 144
    cfg = dict(num_core=64, core_padding=300, M_per_core=50)
    P = np.concatenate([m["positions"][m["indices"].astype(np.int64)].mean(1) for m in sc.meshes]).astype(np.float32)
    N = np.tile(np.array([[0, 1, 0]], np.float32), (P.shape[0], 1))
    trees = random_trees_and_gamma(pkg, P, N, K, KL, lambda p, s, k, b: p.build_tree(s, k, b), seed=9)
    cx, dx = _setup(pkg, False, sc, K, KL, w, h, cfg, trees)
    cf, dfa = _setup(pkg, True, sc, K, KL, w, h, cfg, trees)
    assert cf.fast and not cx.fast

    # (1) traversal: same kernels in both libraries -> bit-equal hits
    rays = pkg.scenes.random_rays(sc, 100000, seed=3)
    hx, hf = cx.trace(rays), cf.trace(rays)
    assert np.array_equal(hx.view(np.uint32), hf.view(np.uint32))
    assert np.array_equal(cx.occlusion(rays), cf.occlusion(rays))

    # (2) one light trace with the exact flavour; both flavours bin and render from the SAME LVC
    dx.P["lt"]["launch_frame"] = 1
    cx.set_params(dx.P)
    cx.launch(pkg.LAUNCH_LIGHT_TRACE, cfg["num_core"], 1)
    cx.synchronize()
    lvc, valid = dx.lvc_host()
    dfa.upload_lvc(lvc, valid)
    out = {}
    for name, ctx, df in (("exact", cx, dx), ("fast", cf, dfa)):
        df.set_sampler_record(ctx.lvc_process(df.lvc, df.valid, df.n_lvc))
        fp = torch.zeros(w * h, dtype=torch.int32, device="cuda")
        fl = torch.zeros(w * h, dtype=torch.int32, device="cuda")
        ctx.set_debug_outputs(fp, fl)
        df.P["subframe_index"] = 0
        df.P["max_depth"] = 1          # camera vertex + one bounce: inputs of every stage are (nearly) identical in both flavours
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_SPCBPT_EYE, w, h)
        ctx.synchronize()
        out[name] = (fp.cpu().numpy(), fl.cpu().numpy(), df.accum.cpu().numpy()[:, :3].copy())
        ctx.set_debug_outputs(None, None)
    # the binning (weights, cmfs) is the same exact-flag code in both libraries
    sx, sf = dx.sampler_host(), dfa.sampler_host()
    assert np.array_equal(sx[2], sf[2]) and np.array_equal(sx[1].view(np.uint32), sf[1].view(np.uint32))
    assert np.array_equal(out["exact"][0], out["fast"][0]), "primary-hit prim ids differ between the flavours"
    lab_diff = (out["exact"][1] != out["fast"][1]).mean()
    assert lab_diff <= 1e-5, "first-vertex subspace ids differ on %.2e of the pixels" % lab_diff
    a, b = out["exact"][2], out["fast"][2]
    lit = a.sum(1) > 1e-4
    rel = np.abs(a[lit] - b[lit]).max(1) / a[lit].max(1)
    print("fast vs exact, depth-1 radiance: median rel %.2e, 90%% %.2e, 99%% %.2e, share > 1e-5: %.4f" % (
        np.median(rel), np.quantile(rel, 0.9), np.quantile(rel, 0.99), (rel > 1e-5).mean()))
    assert np.median(rel) <= 1e-5, "shading differs by more than 1e-5 relative on the median pixel"
    assert np.quantile(rel, 0.9) <= 1e-4
    assert abs(a.mean() / b.mean() - 1) < 5e-3

    # (3) the fast flavour's own light trace: depth-0 / depth-1 vertices of the first cores agree with the exact flavour's
    dfa.P["lt"]["launch_frame"] = 1
    dfa.P["max_depth"] = 0
    cf.set_params(dfa.P)
    cf.launch(pkg.LAUNCH_LIGHT_TRACE, cfg["num_core"], 1)
    cf.synchronize()
    lf, vf = dfa.lvc_host()
    first = np.arange(cfg["num_core"]) * cfg["core_padding"]          # slot 0 of every core: the first emitter sample (same draws)
    assert np.allclose(lf["position"][first], lvc["position"][first], rtol=1e-5, atol=1e-5)
    assert np.array_equal(lf["subspaceId"][first], lvc["subspaceId"][first])
    cx.close()
    cf.close()


def test_fast_flavour_converged_image_matches_exact(gpu_ctx):
    """whole renders (training + 96 frames) with either flavour: same image within Monte-Carlo noise"""
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    kw = dict(K=64, K_light=12, lt_num_core=200, lt_core_padding=400, lt_M_per_core=50, pretrace_num_core=20000)
    imgs = {}
    for fast in (False, True):
        r = Renderer(sc, 128, 96, fast=fast, **kw)
        st = r.preprocessing(target_samples=80000, target_Q_samples=60000, tree_samples=30000, batch_size=20000)
        assert np.isfinite(st["loss_last"])
        for _ in range(256):
            r.render_frame()
        imgs[fast] = r.image().copy()
        r.ctx.close()
    a, b = imgs[False], imgs[True]
    assert np.isfinite(b).all()
    relmse = float(np.mean((a - b) ** 2 / (a ** 2 + 1e-2)))
    print("fast vs exact at 256 spp: means %.5f / %.5f, relMSE between them %.5f" % (a.mean(), b.mean(), relmse))
    assert abs(a.mean() / b.mean() - 1) < 0.01
    assert relmse < 0.01          # two independent 256-spp estimates of the same image (different random decisions after bounce 1)
