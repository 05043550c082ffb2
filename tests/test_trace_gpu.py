"""GPU parity: closest-hit prim ids / t / barycentrics and occlusion bits of the CUDA traversal
(through the C ABI) against the CPU oracle on identical ray batches.  Integer outputs bit-exact;
t,u,v are compared bit-exact too because both sides execute the same IEEE operation sequence
(intersection contract in csrc/traverse.cuh) -- the 1e-5 relative tolerance of the north star is
the fallback stated in the assert message."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _compare(hg, ho, what):
    bad = np.nonzero(hg["prim"] != ho["prim"])[0]
    assert bad.size == 0, "%s: %d prim-id mismatches, first %s gpu=%s oracle=%s" % (what, bad.size, bad[:5], hg[bad[:5]], ho[bad[:5]])
    hit = ho["prim"] >= 0
    for k in ("t", "u", "v"):
        same = hg[k][hit].view(np.uint32) == ho[k][hit].view(np.uint32)
        rel = np.abs(hg[k][hit] - ho[k][hit]) / np.maximum(np.abs(ho[k][hit]), 1e-20)
        assert same.all(), "%s: %s not bit-exact on %d rays (max rel err %g; tolerance 1e-5)" % (what, k, (~same).sum(), rel.max())


@pytest.fixture(scope="module")
def cornell(gpu_ctx, orc):
    pkg = gpu_ctx
    sc = pkg.scenes.cornell_scene()
    ctx = pkg.Context(0, K=64, K_light=12)
    ctx.upload_scene(sc)
    return pkg, sc, ctx, orc.Scene(pkg, sc)


@pytest.fixture(scope="module")
def soup(gpu_ctx, orc):
    pkg = gpu_ctx
    sc = pkg.scenes.random_soup_scene(3000)
    ctx = pkg.Context(0)
    ctx.upload_scene(sc)
    return pkg, sc, ctx, orc.Scene(pkg, sc)


def test_bvh_stats(cornell):
    pkg, sc, ctx, osc = cornell
    st = ctx.bvh_stats()
    assert st["n_triangles"] == sc.n_triangles == 48962
    assert 0 < st["n_nodes"] < st["n_triangles"]
    assert st["max_depth"] <= 32
    assert st["bytes_nodes"] == st["n_nodes"] * 80 and st["bytes_triangles"] == st["n_triangles"] * 48


def test_primary_rays_cornell_512(cornell):
    """config 1: 512x512 primaries, bit-exact prim ids vs the oracle"""
    pkg, sc, ctx, osc = cornell
    rays = pkg.scenes.camera_rays(sc, 512, 512)
    hg = ctx.trace(rays)
    ho = osc.trace(rays, threads=8)
    assert (ho["prim"] >= 0).mean() > 0.5
    _compare(hg, ho, "cornell primaries")


def test_incoherent_rays_cornell(cornell):
    pkg, sc, ctx, osc = cornell
    rays = pkg.scenes.random_rays(sc, 200000, seed=11)
    _compare(ctx.trace(rays), osc.trace(rays, threads=8), "cornell incoherent")
    _compare(ctx.trace(rays, flags=0), osc.trace(rays, flags=0, threads=8), "cornell incoherent no-cull")


def test_soup_vs_bruteforce(soup):
    """adversarial soup (duplicates -> equal-t ties, degenerate + axis-aligned triangles) against the
    brute-force oracle: the result must not depend on the BVH at all"""
    pkg, sc, ctx, osc = soup
    rays = pkg.scenes.random_rays(sc, 20000, seed=5)
    ho = osc.trace(rays, brute=True, threads=8)
    _compare(ctx.trace(rays), ho, "soup")
    cam = pkg.scenes.camera_rays(sc, 128, 128)
    _compare(ctx.trace(cam), osc.trace(cam, brute=True, threads=8), "soup primaries")


def test_axis_aligned_and_edge_rays(cornell):
    """rays along the coordinate axes / lying in wall planes / through shared triangle edges"""
    pkg, sc, ctx, osc = cornell
    r = np.zeros(6 * 4096, pkg.RAY)
    rng = np.random.default_rng(2)
    o = rng.uniform(10, 540, (r.shape[0], 3)).astype(np.float32)
    # snap a third of the origins onto the tessellation grid lines (shared edges)
    o[::3, 0] = np.round(o[::3, 0] / (556.0 / 48)) * np.float32(556.0 / 48)
    d = np.zeros((r.shape[0], 3), np.float32)
    for a in range(6):
        d[a::6, a % 3] = 1.0 if a < 3 else -1.0
    r["ox"], r["oy"], r["oz"] = o.T
    r["dx"], r["dy"], r["dz"] = d.T
    r["tmin"], r["tmax"] = 1e-3, 1e16
    _compare(ctx.trace(r), osc.trace(r, threads=8), "axis rays")


def test_interval_and_culling(cornell):
    """tmin/tmax exclusivity and emitter back-face culling (cuProg.h:402 + Scene.cpp:1030)"""
    pkg, sc, ctx, osc = cornell
    # rays from above the light going down hit its back face: culled -> see the ceiling/ floor instead
    r = np.zeros(4096, pkg.RAY)
    rng = np.random.default_rng(4)
    r["ox"] = rng.uniform(220, 340, r.shape[0])
    r["oz"] = rng.uniform(230, 330, r.shape[0])
    r["oy"] = 548.75
    r["dy"] = -1.0
    r["tmin"], r["tmax"] = 1e-3, 1e16
    hg, ho = ctx.trace(r), osc.trace(r, threads=4)
    _compare(hg, ho, "light back side, cull")
    light_prims = np.arange(sc.n_triangles - 2, sc.n_triangles)
    assert not np.isin(hg["prim"], light_prims).any()
    hg0, ho0 = ctx.trace(r, flags=0), osc.trace(r, flags=0, threads=4)
    _compare(hg0, ho0, "light back side, no cull")
    assert np.isin(hg0["prim"], light_prims).all()
    # shrink tmax to exactly the hit distance: the hit must disappear (t < tmax is strict)
    r2 = r.copy()
    r2["tmax"] = hg0["t"]
    h2 = ctx.trace(r2, flags=0)
    assert (h2["prim"] != hg0["prim"]).all()
    _compare(h2, osc.trace(r2, flags=0, threads=4), "tmax == t")


def test_occlusion(cornell, soup):
    for pkg, sc, ctx, osc in (cornell, soup):
        rng = np.random.default_rng(9)
        P = np.concatenate([m["positions"] for m in sc.meshes])
        lo, hi = P.min(0), P.max(0)
        a = rng.uniform(lo, hi, (100000, 3)).astype(np.float32)
        b = rng.uniform(lo, hi, (100000, 3)).astype(np.float32)
        d = b - a
        ln = np.sqrt((d * d).sum(1)).astype(np.float32)
        d = d / ln[:, None]
        r = np.zeros(a.shape[0], pkg.RAY)
        r["ox"], r["oy"], r["oz"] = a.T
        r["dx"], r["dy"], r["dz"] = d.T
        r["tmin"] = 1e-3
        r["tmax"] = ln - np.float32(1e-3)          # visibilityTest, cuProg.h:466-475
        vg = ctx.occlusion(r)
        vo = osc.occlusion(r, threads=8)
        assert 0.02 < vo.mean() < 0.98
        assert (vg == vo).all(), "%d occlusion mismatches" % (vg != vo).sum()


def test_empty_and_tiny_batches(cornell):
    pkg, sc, ctx, osc = cornell
    assert ctx.trace(np.zeros(0, pkg.RAY)).shape[0] == 0
    assert ctx.occlusion(np.zeros(0, pkg.RAY)).shape[0] == 0
    rays = pkg.scenes.camera_rays(sc, 3, 1)
    _compare(ctx.trace(rays), osc.trace(rays), "3 rays")


def test_single_triangle_scene(gpu_ctx, orc):
    pkg = gpu_ctx
    sc = pkg.scenes.SceneData()
    sc.materials = pkg.scenes.make_pbr(1)
    sc.lights = np.zeros(0, pkg.LIGHT)
    sc.meshes.append(dict(positions=np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32),
                          indices=np.array([[0, 1, 2]], np.uint32), texcoords=None, material_id=0, light_id=-1))
    ctx = pkg.Context(0)
    ctx.upload_scene(sc)
    osc = orc.Scene(pkg, sc)
    r = np.zeros(1000, pkg.RAY)
    rng = np.random.default_rng(1)
    r["ox"], r["oy"] = rng.uniform(-0.5, 1.5, 1000), rng.uniform(-0.5, 1.5, 1000)
    r["oz"] = -1
    r["dz"] = 1
    r["tmin"], r["tmax"] = 1e-3, 1e16
    _compare(ctx.trace(r), osc.trace(r, brute=True), "single triangle")


def test_device_pointer_path_matches_host_path(cornell):
    import torch
    pkg, sc, ctx, osc = cornell
    rays = pkg.scenes.random_rays(sc, 50000, seed=21)
    hh = ctx.trace(rays)
    rd = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    hd = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.trace_device(rd, rays.shape[0], hd)
    torch.cuda.synchronize()
    got = hd.cpu().numpy().view(pkg.HIT).reshape(-1)
    assert (got["prim"] == hh["prim"]).all() and (got["t"].view(np.uint32) == hh["t"].view(np.uint32)).all()
    cnt = ctx.trace_counted(rd, rays.shape[0], hd)
    assert cnt["rays"] == rays.shape[0] and cnt["nodes_visited"] > cnt["rays"] and cnt["tris_tested"] > 0
    ctx.set_stream(0)


def test_scene_share_gives_the_owner_s_results(cornell):
    """spc_scene_share: a second context on the device traces through the first one's BVH (no copy), same hits; bad owners are refused"""
    pkg, sc, ctx, orc_sc = cornell
    other = pkg.Context(0, K=64, K_light=12)
    other.share_scene(ctx)
    assert other.bvh_stats() == ctx.bvh_stats()
    rays = pkg.scenes.camera_rays(sc, 96, 64)
    a, b = ctx.trace(rays), other.trace(rays)
    for k in ("prim", "t", "u", "v"):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), k
    assert np.array_equal(ctx.occlusion(rays), other.occlusion(rays))
    empty = pkg.Context(0, K=64, K_light=12)
    with pytest.raises(pkg.SpcError):
        other.share_scene(empty)          # the owner has no scene
    with pytest.raises(pkg.SpcError):
        other.share_scene(other)          # not itself
    # a borrower that uploads its own scene afterwards stops borrowing (and does not write into the owner's buffers)
    other.upload_scene(pkg.scenes.cornell_scene(wall_cells=2, box_cells=2))
    assert other.bvh_stats()["n_triangles"] != ctx.bvh_stats()["n_triangles"]
    c = ctx.trace(rays)
    assert np.array_equal(a["prim"], c["prim"])
    other.close(); empty.close()


def test_error_paths(gpu_ctx):
    pkg = gpu_ctx
    ctx = pkg.Context(0)
    with pytest.raises(pkg.SpcError):
        ctx.trace(np.zeros(4, pkg.RAY))         # no scene yet -> SPC_ERR_NO_SCENE
    with pytest.raises(pkg.SpcError):
        pkg.Context(0, K=5, K_light=7)          # K_light >= K
