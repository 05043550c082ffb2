"""The reference's OWN device programs on the B200 (oracle/_ref/libref_device.so: raygen.cu + hit_program.cu + cuProg.h + rmis.h
compiled unmodified with --use_fast_math against a device stub <optix.h>; optixTrace = this repository's
traversal) against the product, statistically -- the two run different arithmetic (fast-math intrinsics, CUDA's 9-bit texture
filter) so bits cannot agree, images must:
  * `pt` integrator: 256 spp of the reference program vs 256 spp of ours from the same seeds: means within 0.5 %, per-pixel
    relMSE between the two far below the Monte-Carlo noise of either (same seeds -> mostly the same paths);
  * SPCBPT frame loop (reference light trace + reference LVC_Process + reference eye program, trained state shared): converges
    to the same image as the product's loop.
Skipped where the prebuilt reference libraries are absent (they are built only where /root/reference exists)."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def relmse(a, b):
    return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))


def test_reference_device_programs_agree_with_product(gpu_ctx):
    pkg = gpu_ctx
    rd, rt = _load("ref_device_py"), _load("ref_thrust_py")
    if not (rd.available() and rt.available()):
        pytest.skip("oracle/_ref/libref_device.so / libref_thrust.so not prebuilt")
    from harness import varied_cornell
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(varied_cornell(pkg), 0.01)
    w, h = 160, 120
    # 300 x 400 LVC slots = the size tests/test_ref_thrust_gpu.py uses: the reference's LVC_Process sizes its function-static scratch
    # vectors on the FIRST call of the process (device_thrust.cu:247-250) and overruns them when a later call is larger
    kw = dict(K=1000, lt_num_core=300, lt_core_padding=400, lt_M_per_core=50, pretrace_num_core=20000)
    r = Renderer(sc, w, h, **kw)
    st = r.preprocessing(target_samples=200000, target_Q_samples=100000, tree_samples=50000, batch_size=20000)
    assert np.isfinite(st["loss_last"])
    loop = rd.ReferenceLoop(pkg, r, rt)
    try:
        # pt: same seeds in both (tea<4>(pixel, subframe)), different arithmetic
        for _ in range(256):
            loop.render_frame_pt()
        ref_pt = loop.image().copy()
        ours = Renderer(sc, w, h, **kw)
        for _ in range(256):
            ours.render_frame_pt()
        our_pt = np.nan_to_num(ours.image())
        ref_pt = np.nan_to_num(ref_pt)
        print("pt 256 spp: reference-on-GPU mean %.5f, ours %.5f, relMSE between them %.6f" % (ref_pt.mean(), our_pt.mean(), relmse(our_pt, ref_pt)))
        assert ref_pt.mean() > 0.01 and abs(our_pt.mean() / ref_pt.mean() - 1) < 0.005
        assert relmse(our_pt, ref_pt) < 0.01
        # SPCBPT loop
        loop.subframe = 0
        for _ in range(128):
            loop.render_frame()
        ref_img = loop.image().copy()
        for _ in range(128):
            r.render_frame()
        our_img = r.image()
        print("SPCBPT 128 spp: reference-on-GPU mean %.5f, ours %.5f, relMSE between them %.5f; both vs pt: %.5f / %.5f" % (
            ref_img.mean(), our_img.mean(), relmse(our_img, ref_img), relmse(ref_img, ref_pt), relmse(our_img, ref_pt)))
        assert np.isfinite(ref_img).all() and abs(our_img.mean() / ref_img.mean() - 1) < 0.01
        assert relmse(our_img, ref_img) < 0.02
        assert abs(ref_img.mean() / ref_pt.mean() - 1) < 0.03
    finally:
        loop.close()
