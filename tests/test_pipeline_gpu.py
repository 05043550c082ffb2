"""GPU end-to-end: the reference application's schedule (preprocessing -> per-frame light trace / LVC_Process / eye pass)
driven through the C ABI by the host mirror (spcbpt-optix7_b200/renderer.py).
Config 1 of BASELINE.json: Cornell-class scene (< 100 k triangles), 1 spp per iteration, K = 64 subspaces (12 emitter
subspaces).  The trained estimator must agree with an independent estimate of the same image (the same renderer with
null trees = one subspace, i.e. plain light-vertex-cache connections): relMSE = mean((I-I*)^2 / (I*^2 + 1e-2))."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def test_config1_cornell_spcbpt_vs_pt_ground_truth(gpu_ctx):
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(), 0.01)
    assert sc.n_triangles < 100000
    w = h = 128
    kw = dict(K=64, K_light=12, lt_num_core=200, lt_core_padding=400, lt_M_per_core=50, pretrace_num_core=20000)
    r = Renderer(sc, w, h, **kw)
    st = r.preprocessing(target_samples=120000, target_Q_samples=60000, tree_samples=40000, batch_size=20000)
    assert st["train_paths"] >= 120000 and st["loss_last"] is not None and np.isfinite(st["loss_last"])
    assert len(np.unique(r.eye_tree["label"])) > 32 and r.light_tree["label"].max() < 64 - 12
    for _ in range(64):
        r.render_frame()
    img = r.image().copy()
    # ground truth: the "pt" integrator (an independent estimator: unidirectional + NEE) at 4096 spp
    gt = Renderer(sc, w, h, **kw)
    for _ in range(4096):
        gt.render_frame_pt()
    ref = gt.image()
    assert np.isfinite(img).all() and ref.mean() > 0.02
    # SPCBPT as shipped by the reference drops the t=1 strategy (readme.md:27, rmis.h:137-140) but is otherwise unbiased:
    # image means agree to a few percent and the error falls with the sample count
    assert abs(img.mean() / ref.mean() - 1) < 0.05, (img.mean(), ref.mean())
    e64 = relmse(img, ref)
    pt64 = Renderer(sc, w, h, **kw)
    for _ in range(64):
        pt64.render_frame_pt()
    e_pt64 = relmse(pt64.image(), ref)
    for _ in range(192):
        r.render_frame()
    e256 = relmse(r.image(), ref)
    print("relMSE vs pt@4096spp: SPCBPT 64spp %.5f, 256spp %.5f; pt 64spp %.5f" % (e64, e256, e_pt64))
    assert e64 < 0.05 and e256 < 0.6 * e64, (e64, e256)
    fb = r.frame_rgba8()
    assert fb[..., 3].min() == 255 and fb[..., :3].max() > 100


def test_pipelined_frames_equal_sequential_frames(gpu_ctx):
    """double-buffered LVC + light trace on a side stream: bit-identical images to the sequential loop (same trained state)"""
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    kw = dict(K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)
    r = Renderer(sc, 96, 64, **kw)
    r.preprocessing(target_samples=40000, target_Q_samples=20000, tree_samples=20000, batch_size=20000)
    frame0 = int(r.P["lt"]["launch_frame"][0])
    for _ in range(6):
        r.render_frame()
    a = r.image().copy()
    r.reset_accumulation()
    r.P["lt"]["launch_frame"] = frame0
    r.enable_pipelining()
    for _ in range(6):
        r.render_frame()
    b = r.image().copy()
    assert a.mean() > 0.01 and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_guide_tables_and_compact_trees_leave_frames_bit_identical(gpu_ctx):
    """k_eye_sample's guide tables / compact trees / cached light-vertex labels against the reference's own bisect and tree walk
    (spc_set_option "reference_search" switches them off at launch time): same trained state, same frames, every bit.  K = 64 with 12 emitter
    subspaces on a small LVC leaves several light subspaces empty, so the empty-subspace draw shifts are exercised too."""
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    kw = dict(K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)
    r = Renderer(sc, 96, 64, **kw)
    r.preprocessing(target_samples=40000, target_Q_samples=20000, tree_samples=20000, batch_size=20000)
    frame0 = int(r.P["lt"]["launch_frame"][0])
    assert r.ctx.get_option("reference_search") == 0
    for _ in range(5):
        r.render_frame()
    a = r.image().copy()
    r.reset_accumulation()
    r.P["lt"]["launch_frame"] = frame0
    r.ctx.set_option("reference_search", 1)
    for _ in range(5):
        r.render_frame()
    r.ctx.set_option("reference_search", 0)
    b = r.image().copy()
    assert a.mean() > 0.01 and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_shipped_scene_spcbpt_beats_pt_at_equal_spp(gpu_ctx):
    """BASELINE.json configs[2] in small: the shipped house scene (its .spcscene cache is written by __graft_entry__.build() from
    the reference's data where that exists and travels with the tree; skipped otherwise).  Ground truth = the `pt` integrator at
    2048 spp from disjoint samples; SPCBPT with the reference's full training schedule must be unbiased against it (t=1 strategy
    is dropped by the reference, readme.md:27, which is invisible at this tolerance) and beat `pt` at equal spp."""
    import os
    pkg = gpu_ctx
    cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "_ref", "house.spcscene")
    if not os.path.exists(cache):
        pytest.skip("data/_ref/house.spcscene not present (built only where /root/reference exists)")
    from spcbpt_optix7_b200.renderer import LaneRenderer, Renderer
    sc = pkg.scenes.load_spcscene(cache)
    assert sc.n_triangles == 119140 and len(sc.textures) == 6 and sc.lights.shape[0] == 2
    w, h = 480, 250
    gt = Renderer(sc, w, h, K=1000)
    acc, cnt = np.zeros((h, w, 3)), np.zeros((h, w, 1))
    for c in range(8):                      # chunks: the reference's pt has no NaN guard, a poisoned pixel-chunk is dropped
        gt.reset_accumulation()
        gt.ctx.set_seed_offset(7777777 + 256 * c)
        for _ in range(256):
            gt.render_frame_pt()
        img = gt.image()
        ok = np.isfinite(img).all(-1, keepdims=True)
        acc += np.where(ok, img, 0.0)
        cnt += ok
    ref = (acc / np.maximum(cnt, 1)).astype(np.float32)
    lr = LaneRenderer(sc, w, h, lanes=3, K=1000)
    st = lr.preprocessing()
    assert st["train_paths"] >= 2000000 and np.isfinite(st["loss_last"])
    lr.render(64)
    img = lr.image()
    pt = Renderer(sc, w, h, K=1000)
    for _ in range(64):
        pt.render_frame_pt()
    e_spc, e_pt = relmse(img, ref), relmse(np.nan_to_num(pt.image()), ref)
    print("house 480x250 vs pt@2048spp: relMSE SPCBPT 64spp %.4f, pt 64spp %.4f, means %.4f / %.4f" % (e_spc, e_pt, img.mean(), ref.mean()))
    assert np.isfinite(img).all() and abs(img.mean() / ref.mean() - 1) < 0.02
    assert e_spc < 0.5 * e_pt and e_spc < 0.2


def test_tail_kernel_leaves_frames_bit_identical(gpu_ctx):
    """the eye pass finishes the last few thousand paths in one kernel (spc_set_option "tail_threshold"): whatever the threshold
    -- never, the default, or "as early as the lagged queue read-back allows" -- the frames are the same, every bit, and the
    work counters agree"""
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    kw = dict(K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)
    r = Renderer(sc, 320, 200, **kw)
    r.preprocessing(target_samples=40000, target_Q_samples=20000, tree_samples=20000, batch_size=20000)
    frame0 = int(r.P["lt"]["launch_frame"][0])
    out = []
    for thr in (-1, 0, 4096, 1 << 30):
        r.ctx.set_option("tail_threshold", thr)
        r.reset_accumulation()
        r.P["lt"]["launch_frame"] = frame0
        l0 = r.ctx.launch_count()
        for _ in range(4):
            r.render_frame()
        st = r.ctx.eye_stats()
        out.append((r.image().copy(), st, r.ctx.launch_count() - l0))
    r.ctx.set_option("tail_threshold", 0)
    base = out[0]
    assert base[0].mean() > 0.01
    for img, st, launches in out[1:]:
        assert np.array_equal(base[0].view(np.uint32), img.view(np.uint32))
        assert (st["closest_rays"], st["shadow_rays"], st["visible_connections"]) == (base[1]["closest_rays"], base[1]["shadow_rays"], base[1]["visible_connections"])
    print("launches for 4 frames: no tail %d, default %d, 4096 %d, earliest %d; wavefront bounces %s" % (
        out[0][2], out[1][2], out[2][2], out[3][2], [o[1]["bounces"] for o in out]))
    assert out[3][2] <= out[1][2] < out[0][2] and out[3][2] < out[2][2] < out[0][2] and out[3][1]["bounces"] == 4


def test_sorting_between_bounces_leaves_frames_bit_identical(gpu_ctx):
    """spc_set_option "sort_hits": the wavefront queue re-ordered by hit-point Morton code from the second bounce on -- nothing
    depends on queue order, so frames and work counters are the same, every bit"""
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    kw = dict(K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)
    r = Renderer(sc, 480, 300, **kw)
    r.preprocessing(target_samples=40000, target_Q_samples=20000, tree_samples=20000, batch_size=20000)
    r.ctx.set_option("tail_threshold", 8192)      # keep several wavefront bounces (the sort acts on bounces >= 1 of the wavefront)
    frame0 = int(r.P["lt"]["launch_frame"][0])
    out = []
    for sort in (0, 1):
        r.ctx.set_option("sort_hits", sort)
        r.reset_accumulation()
        r.P["lt"]["launch_frame"] = frame0
        l0 = r.ctx.launch_count()
        for _ in range(4):
            r.render_frame()
        st = r.ctx.eye_stats()
        out.append((r.image().copy(), st, r.ctx.launch_count() - l0))
    r.ctx.set_option("sort_hits", 0)
    r.ctx.set_option("tail_threshold", 0)
    assert out[0][0].mean() > 0.01 and np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
    for k in ("closest_rays", "shadow_rays", "visible_connections", "bounces"):
        assert out[0][1][k] == out[1][1][k], k
    assert out[1][2] > out[0][2]        # the sort kernels did run


@pytest.mark.parametrize("world,dim", [(2, (320, 200)), (3, (203, 77))])
def test_tile_partition_reassembles_the_single_gpu_image(gpu_ctx, world, dim):
    """spc_set_tile_partition (sutil/WorkDistribution.h:34-91 StaticWorkDistribution): every "GPU" renders only the pixels of its
    8 x 4 tiles; ranks that trace the same light paths reproduce, tile by tile, the single-GPU frames bit for bit, the tiles are
    disjoint and cover the image -- also when the image is not a multiple of the strip size"""
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    kw = dict(K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)
    w, h = dim
    r = Renderer(sc, w, h, **kw)
    r.preprocessing(target_samples=40000, target_Q_samples=20000, tree_samples=20000, batch_size=20000)
    frame0 = int(r.P["lt"]["launch_frame"][0])

    def render(gpu_idx, num_gpus):
        r.ctx.set_tile_partition(gpu_idx, num_gpus)
        r.ctx.synchronize()
        r.accum.zero_()
        r.frame.zero_()
        r.reset_accumulation()
        r.P["lt"]["launch_frame"] = frame0
        for _ in range(3):
            r.render_frame()
        r.ctx.synchronize()
        return r.accum.cpu().numpy().reshape(h, w, 4).copy(), r.frame.cpu().numpy().reshape(h, w).copy()

    full, full_frame = render(0, 1)
    assert full[..., :3].mean() > 0.01 and (full[..., 3] == 1.0).all()
    owner = np.full((h, w), -1)
    acc, frm = np.zeros_like(full), np.zeros_like(full_frame)
    for g in range(world):
        a, f = render(g, world)
        mine = a[..., 3] == 1.0
        assert (owner[mine] == -1).all(), "tiles of two GPUs overlap"
        owner[mine] = g
        assert (a[~mine] == 0).all() and (f[~mine] == 0).all(), "a rank wrote outside its tiles"
        acc += a
        frm += f
    r.ctx.set_tile_partition(0, 1)
    assert (owner >= 0).all(), "tiles do not cover the image"
    # the reference's layout: pixel (x, y) belongs to GPU ((x / 8) - (y / 4)) mod world   (WorkDistribution.h:67-80)
    yy, xx = np.mgrid[0:h, 0:w]
    assert np.array_equal(owner, (xx // 8 - yy // 4) % world)
    assert np.array_equal(acc.view(np.uint32), full.view(np.uint32)) and np.array_equal(frm, full_frame)
    with pytest.raises(Exception):
        r.ctx.set_tile_partition(2, 2)
