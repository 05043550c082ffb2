"""GPU parity on the configurations the numbers are quoted on: the whole per-frame chain
(light trace -> LVC_Process -> eye pass) through the C ABI against the CPU oracle, bit for bit, on
  * the textured + metallic + two-light Cornell fixture (row a5: ColorTexSample / sampleTexture / powf(c, 2.2),
    hit_program.cu:182-198, src/cuda/LocalShading.h:37-53) -- the same fixture tests/test_oracle_vs_ref.py pins to the
    reference's own programs on the host;
  * BASELINE.json configs[0]: Cornell 48 962 triangles at 512x512, K = 64 subspaces (12 emitter subspaces): primary-hit
    prim ids and first-vertex subspace ids exact, accumulation buffer bit-equal;
  * the shipped house scene (119 140 triangles, 6 textures, 2 lights) at low resolution, K = 1000;
  * BASELINE.json configs[1]: the 1 002 542-triangle height field -- closest hits and occlusion of the bench's own ray
    sets (primary, cosine bounce, shadow) on a 2^18-ray sample.
Trees come from the product's host builder (== the reference's builder, tests/test_oracle_vs_ref.py) over the triangle
centroids; Q / Gamma are seeded random tables (harness.random_q_gamma)."""
import os

import numpy as np
import pytest

from harness import DeviceFrame, HostFrame, compare_lvc, float_bits_differ, random_trees_and_gamma, varied_cornell

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _centroids(sc, cap=20000, seed=0):
    P, N = [], []
    for m in sc.meshes:
        tri = m["positions"][m["indices"].astype(np.int64)]
        n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
        P.append(tri.mean(1).astype(np.float32))
        N.append(n.astype(np.float32))
    P, N = np.concatenate(P), np.concatenate(N)
    if P.shape[0] > cap:
        sel = np.sort(np.random.default_rng(seed).choice(P.shape[0], cap, replace=False))
        P, N = P[sel], N[sel]
    return P, N


def _chain_vs_oracle(pkg, orc, sc, K, KL, w, h, cfg, subframes, max_bad, seed, max_depth=0, connections=3, options=()):
    """light trace -> LVC_Process -> eye pass on GPU and oracle; returns (first_prim, first_label, accum) of the last subframe"""
    import torch
    P, N = _centroids(sc)
    eye_tree, light_tree, Q, cmf = random_trees_and_gamma(pkg, P, N, K, KL, lambda p, s, k, b: p.build_tree(s, k, b), seed=seed)
    ctx = pkg.Context(0, K=K, K_light=KL, connections=connections)
    ctx.upload_scene(sc)
    for name, value in options:
        ctx.set_option(name, value)
    df = DeviceFrame(pkg, sc, w, h, K=K, **cfg)
    hf = HostFrame(pkg, sc, w, h, K=K, **cfg)
    # trees go through spc_tree_to_device (compact copies, cross labels): the production path of the eye pass
    df.P["subspace_info"]["eye_tree"] = ctx.tree_to_device(True, eye_tree)
    df.P["subspace_info"]["light_tree"] = ctx.tree_to_device(False, light_tree)
    hf.set_trees(eye_tree, light_tree)
    df.P["max_depth"] = hf.P["max_depth"] = max_depth      # MyParams::max_depth (0: the reference's literal 50)
    df.set_q_gamma(Q, cmf)   # a caller-made CDF: searched with the reference's bisect (guide tables exist only for tables the library built)
    hf.set_q_gamma(Q, cmf)
    osc = orc.Scene(pkg, sc)
    fp = torch.zeros(w * h, dtype=torch.int32, device="cuda")
    fl = torch.zeros(w * h, dtype=torch.int32, device="cuda")
    ctx.set_debug_outputs(fp, fl)
    out = None
    for sf in subframes:
        for f in (df, hf):
            f.P["subframe_index"] = sf
            f.P["lt"]["launch_frame"] = sf + 1
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_LIGHT_TRACE, cfg["num_core"], 1)
        df.set_sampler_record(ctx.lvc_process(df.lvc, df.valid, df.n_lvc))
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_SPCBPT_EYE, w, h)
        ctx.synchronize()
        orc.light_trace(osc, hf.P, K, max_depth=max_depth, threads=8, connections=connections)
        sub, cmfs, jump, vc, pc = orc.lvc_process(pkg, hf.lvc, hf.valid, K)
        hf.set_sampler(sub, cmfs, jump, vc, pc)
        ofp, ofl = orc.eye_pass(osc, hf.P, K, connections, max_depth, threads=8, want_first=True)
        lvc, valid = df.lvc_host()
        badv = compare_lvc(pkg, lvc, valid, hf.lvc, hf.valid, exact=True)
        assert not badv, "subframe %d LVC: %s" % (sf, badv)
        gs, gc, gj, gvc, gpc = df.sampler_host()
        assert (gvc, gpc) == (vc, pc) and np.array_equal(gj, jump) and not float_bits_differ(gc, cmfs).any()
        gfp, gfl = fp.cpu().numpy(), fl.cpu().numpy()
        assert np.array_equal(gfp, ofp), "subframe %d: %d primary-hit prim ids differ" % (sf, int((gfp != ofp).sum()))
        assert np.array_equal(gfl, ofl), "subframe %d: %d first-vertex subspace ids differ" % (sf, int((gfl != ofl).sum()))
        acc = df.accum.cpu().numpy()
        bad = float_bits_differ(acc, hf.accum).any(1)
        assert bad.sum() <= max_bad, "subframe %d: %d of %d pixels differ; first %s gpu %s oracle %s" % (
            sf, bad.sum(), bad.size, np.nonzero(bad)[0][:3], acc[bad][:3], hf.accum[bad][:3])
        assert np.array_equal(df.frame.cpu().numpy().view(np.uint32)[~bad], hf.frame[~bad])
        out = (gfp, gfl, acc, vc)
    ctx.set_debug_outputs(None, None)
    ctx.close()
    return out


def test_textured_metallic_two_light_scene_bit_exact(gpu_ctx, orc):
    """row a5: the texture fetch + powf(c, 2.2) linearisation, the metallic lobe and the two-light emitter pick on the GPU"""
    pkg = gpu_ctx
    sc = varied_cornell(pkg)
    cfg = dict(num_core=48, core_padding=200, M_per_core=40)
    fp, fl, acc, vc = _chain_vs_oracle(pkg, orc, sc, 1000, 200, 128, 96, cfg, (0, 1, 5), 2, seed=9)
    # the textured box (mesh 4) and the metal box (mesh 3) are in view and lit
    first = np.cumsum([0] + [m["indices"].shape[0] for m in sc.meshes])
    on_tex = (fp >= first[4]) & (fp < first[5])
    on_metal = (fp >= first[3]) & (fp < first[4])
    assert on_tex.sum() > 300 and on_metal.sum() > 300, (on_tex.sum(), on_metal.sum())
    assert acc[on_tex, :3].mean() > 0.005 and acc[on_metal, :3].mean() > 0.005
    assert vc > 1500


def test_config1_cornell_512_K64_ids_bit_exact(gpu_ctx, orc):
    """BASELINE.json configs[0]: 512x512, K = 64, 12 emitter subspaces -- prim ids + subspace ids exact, accum bit-equal"""
    pkg = gpu_ctx
    sc = pkg.scenes.cornell_scene()
    assert sc.n_triangles == 48962
    cfg = dict(num_core=100, core_padding=400, M_per_core=50)
    fp, fl, acc, vc = _chain_vs_oracle(pkg, orc, sc, 64, 12, 512, 512, cfg, (0, 1), 4, seed=4)
    assert (fp >= 0).mean() > 0.9 and len(np.unique(fl[fl >= 0])) > 20 and fl.max() < 64


def test_house_scene_low_res_bit_exact(gpu_ctx, orc):
    """the shipped scene (textures, 119 140 triangles, two divLevel-10 lights), K = 1000, 320x180"""
    pkg = gpu_ctx
    cache = os.path.join(ROOT, "data", "_ref", "house.spcscene")
    if not os.path.exists(cache):
        pytest.skip("data/_ref/house.spcscene not present (built only where /root/reference exists)")
    sc = pkg.scenes.load_spcscene(cache)
    cfg = dict(num_core=200, core_padding=800, M_per_core=100)
    fp, fl, acc, vc = _chain_vs_oracle(pkg, orc, sc, 1000, 200, 320, 180, cfg, (0, 1), 4, seed=6)
    assert (fp >= 0).mean() > 0.9 and len(np.unique(fl[fl >= 0])) > 50 and vc > 30000
    assert np.isfinite(acc).all() and acc[:, :3].mean() > 0.05


def test_config5_class_glossy_many_emitters_depth12_bit_exact(gpu_ctx, orc):
    """BASELINE.json configs[4] in miniature: the large-scene generator at 96 x 96 quads (18 432 terrain triangles, rough-metal and
    glossy materials, 16 quad emitters with one subspace each), max depth 12, K = 80 with 16 emitter subspaces"""
    pkg = gpu_ctx
    sc = pkg.scenes.large_scene(96, 4)
    assert sc.lights.shape[0] == 16
    cfg = dict(num_core=64, core_padding=400, M_per_core=40)
    fp, fl, acc, vc = _chain_vs_oracle(pkg, orc, sc, 80, 16, 256, 144, cfg, (0, 1), 4, seed=12, max_depth=12)
    assert (fp >= 0).mean() > 0.9 and vc > 3000 and np.isfinite(acc).all() and acc[:, :3].mean() > 1e-3


@pytest.mark.parametrize("connections,options", [(1, ()), (2, ()), (4, ()), (6, ()), (6, (("tail_threshold", 1 << 30),)), (2, (("tail_threshold", -1),))])
def test_other_connection_counts_bit_exact(gpu_ctx, orc, connections, options):
    """CONNECTION_N is a runtime value here (optixPathTracer.h:33 fixes it at 3): the specialised kernels for 1, 2 and 4 connections and
    the generic ones (> 4), through the wavefront stages and through the tail kernel"""
    pkg = gpu_ctx
    sc = pkg.scenes.cornell_scene(wall_cells=6, box_cells=4)
    cfg = dict(num_core=32, core_padding=160, M_per_core=30)
    fp, fl, acc, vc = _chain_vs_oracle(pkg, orc, sc, 64, 12, 96, 64, cfg, (0, 1), 2, seed=5, connections=connections, options=options)
    assert np.isfinite(acc).all() and acc[:, :3].mean() > 0.01


def test_light_vertex_cache_overflow_bit_exact(gpu_ctx, orc):
    """a core whose window of `core_padding` slots fills up before its M_per_core paths are traced (raygen.cu:640-684: the path in flight
    is cut, the remaining paths are skipped): same LVC, same sampler, same frames"""
    pkg = gpu_ctx
    sc = pkg.scenes.cornell_scene(wall_cells=6, box_cells=4)
    cfg = dict(num_core=24, core_padding=24, M_per_core=20)
    fp, fl, acc, vc = _chain_vs_oracle(pkg, orc, sc, 64, 12, 64, 48, cfg, (0, 1), 2, seed=8)
    assert vc > 24 * 20 and np.isfinite(acc).all()      # the windows are (nearly) full


@pytest.mark.parametrize("dim", [(1, 1), (7, 5), (33, 1), (1, 40)])
def test_tiny_and_odd_image_sizes_bit_exact(gpu_ctx, orc, dim):
    """edge sizes of the launch (a single pixel, fewer pixels than a warp, one row, one column): queues, compaction and the tail kernel
    at their boundaries"""
    pkg = gpu_ctx
    sc = pkg.scenes.cornell_scene(wall_cells=6, box_cells=4)
    cfg = dict(num_core=16, core_padding=120, M_per_core=20)
    fp, fl, acc, vc = _chain_vs_oracle(pkg, orc, sc, 64, 12, dim[0], dim[1], cfg, (0, 1, 2), 0, seed=3)
    assert fp.shape[0] == dim[0] * dim[1] and np.isfinite(acc).all()


def test_config2_heightfield_bench_ray_sets_vs_oracle(gpu_ctx, orc):
    """BASELINE.json configs[1]: the bench's own three ray sets on the 1 M-triangle height field, 2^18-ray sample each:
    prim ids, t/u/v bits and visibility against the oracle (bench.py repeats this on the full 2^24-ray sets)"""
    import torch
    pkg = gpu_ctx
    sc = pkg.scenes.heightfield_scene(708)
    assert sc.n_triangles == 1002542
    ctx = pkg.Context(0)
    ctx.upload_scene(sc)
    side = 512
    n = side * side
    eye, U, V, W = sc.camera_frame(side, side)
    cam = np.concatenate([eye, U, V, W]).astype(np.float32)
    A = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    B, C = torch.empty_like(A), torch.empty_like(A)
    hA = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    hB = torch.empty_like(hA)
    vC = torch.empty(n, dtype=torch.uint8, device="cuda")
    ctx.call("spc_gen_camera_rays", cam, side, side, 1, A)
    ctx.trace_device(A, n, hA)
    ctx.call("spc_gen_bench_rays", 1, A, hA, n, B, None)
    ctx.call("spc_gen_bench_rays", 2, A, hA, n, C, None)
    ctx.trace_device(B, n, hB)
    ctx.occlusion_device(C, n, vC)
    ctx.synchronize()
    osc = orc.Scene(pkg, sc)
    for name, rays, hits in (("A", A, hA), ("B", B, hB)):
        ho = osc.trace(rays.cpu().numpy().view(pkg.RAY).reshape(-1), threads=8)
        hg = hits.cpu().numpy().view(pkg.HIT).reshape(-1)
        assert np.array_equal(hg["prim"], ho["prim"]), "set %s: %d prim ids differ" % (name, int((hg["prim"] != ho["prim"]).sum()))
        for k in ("t", "u", "v"):
            assert not float_bits_differ(hg[k], ho[k]).any(), "set %s: %s not bit-exact" % (name, k)
        assert (ho["prim"] >= 0).mean() > 0.9
    vo = osc.occlusion(C.cpu().numpy().view(pkg.RAY).reshape(-1), threads=8)
    assert np.array_equal(vC.cpu().numpy(), vo)
    assert 0.05 < vo.mean() < 0.95
    ctx.close()
