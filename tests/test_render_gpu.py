"""GPU parity of the render path (light trace -> LVC binning -> eye pass), through the C ABI, against
 (a) the committed golden vectors produced by the reference's own programs (tests/golden/render.npz) and
 (b) the CPU oracle on larger seeded inputs.
Kernels and oracle follow one arithmetic policy (DESIGN.md "bit parity"), so integer outputs AND floats are
compared bit-for-bit; `MAX_BAD` bounds the pixels/vertices allowed to differ through the one documented
source of divergence (fp64 libm results that straddle an fp32 rounding boundary, p ~ 1e-8 per call)."""
import hashlib
import os

import numpy as np
import pytest

from harness import GOLDEN_CFG, DeviceFrame, HostFrame, compare_lvc, float_bits_differ, golden_scene, random_q_gamma

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAX_BAD = 2


@pytest.fixture(scope="module")
def gold(gpu_ctx, orc):
    pkg = gpu_ctx
    g = np.load(os.path.join(GOLD, "render.npz"))
    sc = golden_scene(pkg)
    K = 1000
    Q, cmf = random_q_gamma(K, 1000)
    assert hashlib.sha256(cmf.tobytes()).hexdigest() == str(g["cmf_sha"])
    ctx = pkg.Context(0, K=K, K_light=200, connections=3)
    ctx.upload_scene(sc)
    c = GOLDEN_CFG
    df = DeviceFrame(pkg, sc, c["w"], c["h"], K=K, num_core=c["num_core"], core_padding=c["core_padding"], M_per_core=c["M_per_core"])
    df.set_trees(g["eye_tree"], g["light_tree"])
    df.set_q_gamma(Q, cmf)
    df.P["lt"]["launch_frame"] = c["launch_frame"]
    return pkg, g, sc, ctx, df, K, Q, cmf


def test_light_trace_vs_reference_golden(gold):
    """LVC of one light-trace launch == the LVC the reference's own __raygen__lightTrace produced"""
    pkg, g, sc, ctx, df, K, Q, cmf = gold
    ctx.set_params(df.P)
    ctx.launch("light trace", GOLDEN_CFG["num_core"], 1)
    ctx.synchronize()
    lvc, valid = df.lvc_host()
    bad = compare_lvc(pkg, lvc, valid, g["lvc"], g["valid"], exact=True)
    assert not bad, bad
    assert int(valid.sum()) == int(g["vc"]) > 500


def test_lvc_process_vs_oracle(gold, orc):
    pkg, g, sc, ctx, df, K, Q, cmf = gold
    df.upload_lvc(g["lvc"], g["valid"])
    rec = ctx.lvc_process(df.lvc, df.valid, df.n_lvc)
    df.set_sampler_record(rec)
    sub, cmfs, jump, vc, pc = df.sampler_host()
    assert vc == int(g["vc"]) and pc == int(g["pc"])
    assert np.array_equal(jump, g["jump"])
    d = float_bits_differ(cmfs, g["cmfs"])
    assert not d.any(), "%d cmf entries differ, first at %s: gpu %s oracle %s" % (d.sum(), np.nonzero(d)[0][:4], cmfs[d][:4], g["cmfs"][d][:4])
    for k in ("jump_bias", "id", "size"):
        assert np.array_equal(sub[k], g["sub"][k]), k
    assert not float_bits_differ(sub["sum_pmf"], g["sub"]["sum_pmf"]).any()


def test_eye_pass_vs_reference_golden_and_oracle(gold, orc):
    pkg, g, sc, ctx, df, K, Q, cmf = gold
    df.upload_lvc(g["lvc"], g["valid"])
    df.set_sampler_record(ctx.lvc_process(df.lvc, df.valid, df.n_lvc))
    # oracle twin on the host, nvcc's left-to-right jitter order
    c = GOLDEN_CFG
    hf = HostFrame(pkg, sc, c["w"], c["h"], K=K, num_core=c["num_core"], core_padding=c["core_padding"], M_per_core=c["M_per_core"])
    hf.set_trees(g["eye_tree"], g["light_tree"])
    hf.set_q_gamma(Q, cmf)
    hf.lvc[:] = g["lvc"]
    hf.valid[:] = g["valid"]
    hf.set_sampler(g["sub"].copy(), g["cmfs"].copy(), g["jump"].copy(), int(g["vc"]), int(g["pc"]))
    osc = orc.Scene(pkg, sc)
    import torch
    fp = torch.zeros(c["w"] * c["h"], dtype=torch.int32, device="cuda")
    fl = torch.zeros(c["w"] * c["h"], dtype=torch.int32, device="cuda")
    ctx.set_debug_outputs(fp, fl)
    for k, sf in enumerate(c["subframes"]):
        df.P["subframe_index"] = sf
        hf.P["subframe_index"] = sf
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_SPCBPT_EYE, c["w"], c["h"])
        ctx.synchronize()
        ofp, ofl = orc.eye_pass(osc, hf.P, K, 3, 0, threads=8, want_first=True)
        acc = df.accum.cpu().numpy()
        assert np.array_equal(fp.cpu().numpy(), ofp), "primary-hit prim ids, subframe %d" % sf
        assert np.array_equal(fl.cpu().numpy(), ofl), "first-vertex subspace ids, subframe %d" % sf
        bad = float_bits_differ(acc, hf.accum).any(1)
        assert bad.sum() <= MAX_BAD, "subframe %d: %d pixels differ from the oracle; first %s gpu %s oracle %s" % (
            sf, bad.sum(), np.nonzero(bad)[0][:3], acc[bad][:3], hf.accum[bad][:3])
        assert np.array_equal(df.frame.cpu().numpy().view(np.uint32)[~bad], hf.frame[~bad])
        if sf == 0:   # no jitter at subframe 0: the reference's own accum buffer is directly comparable
            badr = float_bits_differ(acc, g["accum"][0]).any(1)
            assert badr.sum() <= MAX_BAD, "%d pixels differ from the reference golden" % badr.sum()
    ctx.set_debug_outputs(None, None)
    assert acc[:, :3].mean() > 0.01


def test_full_frame_pipeline_vs_oracle(gpu_ctx, orc):
    """config 1 class: Cornell 49 k triangles, 256x256, K=1000, the whole per-frame chain on the GPU (light trace ->
    LVC_Process -> eye pass, 2 subframes) against the oracle running the same chain on the host"""
    pkg = gpu_ctx
    g = np.load(os.path.join(GOLD, "render.npz"))
    sc = pkg.scenes.cornell_scene()
    K = 1000
    Q, cmf = random_q_gamma(K, 1000)
    ctx = pkg.Context(0, K=K, K_light=200, connections=3)
    ctx.upload_scene(sc)
    w = h = 256
    cfg = dict(num_core=200, core_padding=400, M_per_core=50)
    df = DeviceFrame(pkg, sc, w, h, K=K, **cfg)
    hf = HostFrame(pkg, sc, w, h, K=K, **cfg)
    for f in (df, hf):
        f.set_trees(g["eye_tree"], g["light_tree"])
        f.set_q_gamma(Q, cmf)
    osc = orc.Scene(pkg, sc)
    for sf in (0, 1):
        for f in (df, hf):
            f.P["subframe_index"] = sf
            f.P["lt"]["launch_frame"] = sf + 1
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_LIGHT_TRACE, cfg["num_core"], 1)
        df.set_sampler_record(ctx.lvc_process(df.lvc, df.valid, df.n_lvc))
        ctx.set_params(df.P)
        ctx.launch(pkg.LAUNCH_SPCBPT_EYE, w, h)
        ctx.synchronize()
        orc.light_trace(osc, hf.P, K, threads=8)
        sub, cmfs, jump, vc, pc = orc.lvc_process(pkg, hf.lvc, hf.valid, K)
        hf.set_sampler(sub, cmfs, jump, vc, pc)
        orc.eye_pass(osc, hf.P, K, 3, 0, threads=8)
        lvc, valid = df.lvc_host()
        badv = compare_lvc(pkg, lvc, valid, hf.lvc, hf.valid, exact=True)
        assert not badv, badv
        gs, gc, gj, gvc, gpc = df.sampler_host()
        assert (gvc, gpc) == (vc, pc) and np.array_equal(gj, jump) and not float_bits_differ(gc, cmfs).any()
        acc = df.accum.cpu().numpy()
        bad = float_bits_differ(acc, hf.accum).any(1)
        assert bad.sum() <= 4, "subframe %d: %d of %d pixels differ" % (sf, bad.sum(), bad.size)
    assert vc > 20000


def test_pt_integrator_vs_oracle(gold, orc):
    """the "pt" comparison integrator (raygen.cu:71-170) against its oracle restatement (pinned to the reference's own program
    in tests/test_oracle_vs_ref.py): four subframes, bit-exact accumulation"""
    pkg, g, sc, ctx, df, K, Q, cmf = gold
    c = GOLDEN_CFG
    hf = HostFrame(pkg, sc, c["w"], c["h"], K=K, num_core=c["num_core"], core_padding=c["core_padding"], M_per_core=c["M_per_core"])
    osc = orc.Scene(pkg, sc)
    df.accum.zero_()
    for sf in (0, 1, 2, 3):
        df.P["subframe_index"] = sf
        hf.P["subframe_index"] = sf
        ctx.set_params(df.P)
        ctx.launch("pt", c["w"], c["h"])
        ctx.synchronize()
        orc.pt_pass(osc, hf.P, K, threads=8)
        bad = float_bits_differ(df.accum.cpu().numpy(), hf.accum).any(1)
        assert bad.sum() <= MAX_BAD, "subframe %d: %d pixels differ" % (sf, bad.sum())
    assert hf.accum[:, :3].mean() > 0.01
