"""GPU: the C++ host driver (host/spcbpt_main.cpp = the reference application's schedule above the C ABI) against the
Python mirror of the same schedule (spcbpt-optix7_b200/renderer.py).  Same scene files, same launches, same seeds ->
the accumulation buffers must agree bit for bit, through the .scene/OBJ loader, the .spcscene cache and the pipelined loop."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "_build", "spcbpt_render")


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = (int(v) for v in f.readline().split())
        assert float(f.readline()) < 0
        return np.frombuffer(f.read(), "<f4").reshape(h, w, 3)


def read_ppm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"P6"
        w, h = (int(v) for v in f.readline().split())
        f.readline()
        return np.frombuffer(f.read(), np.uint8).reshape(h, w, 3)


def run_driver(*args):
    assert os.path.exists(BIN), "host/_build/spcbpt_render missing: run __graft_entry__.build()"
    r = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1]), r.stdout


SMALL = ["--K", 64, "--K-light", 12, "--lt-cores", 100, "--lt-padding", 300, "--lt-per-core", 40, "--pretrace-cores", 20000,
         "--train-samples", 40000, "--q-samples", 20000, "--tree-samples", 20000, "--batch", 20000]


def make_renderer(pkg, sc, w, h):
    from spcbpt_optix7_b200.renderer import Renderer
    return Renderer(sc, w, h, K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)


def test_cpp_driver_equals_python_mirror(gpu_ctx, tmp_path):
    pkg = gpu_ctx
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    path = pkg.scenes.export_scene(sc, str(tmp_path), "cb")
    w, h, frames = 96, 64, 5
    cache = tmp_path / "cb.spcscene"
    dim = "--dim=%dx%d" % (w, h)

    # 1. "pt" needs no training: .scene + OBJ through the C++ loader and driver == the Python mirror on the in-memory scene
    st, _ = run_driver("--scene", path, dim, "--frames", 3, "--alg", "pt", "--out", tmp_path / "c", "--save-cache", cache, *SMALL)
    assert st["triangles"] == sc.n_triangles and st["frames"] == 3 and st["kernel_launches"] > 0
    r = make_renderer(pkg, sc, w, h)
    for _ in range(3):
        r.render_frame_pt()
    img_pt = r.image().copy()
    assert img_pt.mean() > 0.01
    assert np.array_equal(read_pfm(tmp_path / "c.pfm").view(np.uint32), img_pt.view(np.uint32)), "pt: C++ driver and Python mirror disagree"
    assert np.array_equal(read_ppm(tmp_path / "c.ppm"), r.frame_rgba8()[::-1, :, :3])   # PPM = tone-mapped frame buffer, top-down

    # 2. SPCBPT.  Training is tolerance-level by nature (fp32 atomics in the Gamma histogram), so the trained
    #    state is shared through the reference's text files: Python trains and saves, the C++ driver loads -> bit-identical frames
    sc2 = pkg.scenes.load_spcscene(str(cache))
    r = make_renderer(pkg, sc2, w, h)
    r.preprocessing(target_samples=40000, target_Q_samples=20000, tree_samples=20000, batch_size=20000)
    prefix = str(tmp_path / "st_")
    r.save_state(prefix)
    r.P["lt"]["launch_frame"] = 0      # the driver starts its light-trace counter at 0 when it loads the state instead of training
    for _ in range(frames):
        r.render_frame()
    img_py = r.image().copy()
    st1, _ = run_driver("--cache", cache, dim, "--frames", frames, "--out", tmp_path / "a", "--load-state", prefix, "--no-pipeline", *SMALL)
    assert not st1["pipelined"]
    img_cpp = read_pfm(tmp_path / "a.pfm")
    assert np.array_equal(img_cpp.view(np.uint32), img_py.view(np.uint32)), "SPCBPT: C++ driver and Python mirror disagree"
    # pipelined loop (light trace of frame f+1 under the eye pass of frame f): same image
    st2, _ = run_driver("--cache", cache, dim, "--frames", frames, "--out", tmp_path / "b", "--load-state", prefix, *SMALL)
    assert st2["pipelined"]
    assert np.array_equal(read_pfm(tmp_path / "b.pfm").view(np.uint32), img_cpp.view(np.uint32))
    # frame lanes (3 contexts / streams / host threads rendering alternate subframes): the same samples, running means merged
    # at read-out -> equal to the sequential image up to the fp32 summation order
    st4, _ = run_driver("--cache", cache, dim, "--frames", 7, "--out", tmp_path / "l3", "--load-state", prefix, "--lanes", 3, *SMALL)
    st5, _ = run_driver("--cache", cache, dim, "--frames", 7, "--out", tmp_path / "l1", "--load-state", prefix, "--no-pipeline", *SMALL)
    assert st4["lanes"] == 3 and st5["lanes"] == 1
    a3, a1 = read_pfm(tmp_path / "l3.pfm"), read_pfm(tmp_path / "l1.pfm")
    assert np.allclose(a3, a1, rtol=2e-6, atol=1e-7) and not np.array_equal(a3, np.zeros_like(a3))
    assert np.abs(read_ppm(tmp_path / "l3.ppm").astype(int) - read_ppm(tmp_path / "l1.ppm").astype(int)).max() <= 1
    # state files round-trip through the Python loader as well
    r2 = make_renderer(pkg, sc2, w, h)
    r2.load_state(prefix)
    for _ in range(frames):
        r2.render_frame()
    assert np.array_equal(r2.image().view(np.uint32), img_py.view(np.uint32))

    # 3. the driver's own training, no shared state: training has no floating-point atomics (train.cu sums every scatter-add in a
    #    fixed order), so the whole pipeline -- pretrace, trees, Q, Gamma, Adam, render -- is bit-reproducible across processes
    st3, _ = run_driver("--cache", cache, dim, "--frames", frames, "--out", tmp_path / "d", "--save-state", tmp_path / "own_", "--no-pipeline", *SMALL)
    assert st3["train_paths"] >= 40000 and os.path.exists(tmp_path / "own_E.txt")
    r3 = make_renderer(pkg, sc2, w, h)
    r3.preprocessing(target_samples=40000, target_Q_samples=20000, tree_samples=20000, batch_size=20000)
    for _ in range(frames):
        r3.render_frame()
    assert np.array_equal(read_pfm(tmp_path / "d.pfm").view(np.uint32), r3.image().view(np.uint32)), "own training: C++ driver and Python mirror disagree"


def test_python_lane_renderer_equals_sequential(gpu_ctx, tmp_path):
    """LaneRenderer (bench.py's render loop): lanes render the sequential loop's subframes, one thread per lane"""
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import LaneRenderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    w, h = 80, 60
    kw = dict(K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)
    lr = LaneRenderer(sc, w, h, lanes=3, **kw)
    lr.preprocessing(target_samples=40000, target_Q_samples=20000, tree_samples=20000, batch_size=20000)
    prefix = str(tmp_path / "st_")
    lr.lanes[0].save_state(prefix)
    lr.render(4)
    lr.render(3)          # continues where the first call stopped
    assert lr.frames == 7 and lr.lane_counts() == [3, 2, 2]
    img = lr.image().copy()
    seq = make_renderer(pkg, sc, w, h)
    seq.load_state(prefix)
    seq.P["lt"]["launch_frame"] = lr.lt_base
    for _ in range(7):
        seq.render_frame()
    ref = seq.image()
    assert ref.mean() > 0.01 and np.allclose(img, ref, rtol=2e-6, atol=1e-7)
    # each lane's running mean is exactly the mean of its own subframes: lane 1 = sequential frames 1 and 4
    assert np.abs(lr.frame_rgba8().astype(int) - seq.frame_rgba8().astype(int)).max() <= 1


def test_cpp_driver_writes_png(gpu_ctx, tmp_path):
    """<out>.png (host/png_decode.cpp write_png_from_uchar4) holds the pixels of <out>.ppm"""
    from PIL import Image
    pkg = gpu_ctx
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=6, box_cells=4), 0.01)
    path = pkg.scenes.export_scene(sc, str(tmp_path), "cb")
    run_driver("--scene", path, "--dim=64x48", "--frames", 2, "--alg", "pt", "--out", tmp_path / "p", *SMALL)
    assert np.array_equal(np.asarray(Image.open(tmp_path / "p.png").convert("RGB")), read_ppm(tmp_path / "p.ppm"))


def test_cpp_driver_tile_partition_two_ranks(gpu_ctx, tmp_path):
    """--ranks 2 --tiles: the two GPUs render the same subframes, each the pixels of its 8 x 4 tiles (spc_set_tile_partition,
    sutil/WorkDistribution.h:34-91); the NCCL read-out reassembles exactly the single-GPU image"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    pkg = gpu_ctx
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    path = pkg.scenes.export_scene(sc, str(tmp_path), "cb")
    dim, frames = "--dim=100x70", 4
    run_driver("--scene", path, dim, "--frames", 1, "--out", tmp_path / "t", "--save-state", tmp_path / "st_", "--no-images", *SMALL)
    one, _ = run_driver("--scene", path, dim, "--frames", frames, "--out", tmp_path / "one", "--load-state", tmp_path / "st_", *SMALL)
    two, _ = run_driver("--scene", path, dim, "--frames", frames, "--out", tmp_path / "two", "--load-state", tmp_path / "st_", "--ranks", 2, "--tiles", *SMALL)
    assert two["ranks"] == 2 and one["image_mean"] > 0.01
    assert np.array_equal(read_pfm(tmp_path / "one.pfm").view(np.uint32), read_pfm(tmp_path / "two.pfm").view(np.uint32))
    assert np.array_equal(read_ppm(tmp_path / "one.ppm"), read_ppm(tmp_path / "two.ppm"))
    # with frame lanes on each rank as well
    two4, _ = run_driver("--scene", path, dim, "--frames", frames, "--out", tmp_path / "two4", "--load-state", tmp_path / "st_", "--ranks", 2, "--tiles", "--lanes", 2, *SMALL)
    assert np.allclose(read_pfm(tmp_path / "two4.pfm"), read_pfm(tmp_path / "one.pfm"), rtol=2e-6, atol=1e-7)


def test_cpp_driver_errors(gpu_ctx, tmp_path):
    r = subprocess.run([BIN, "--scene", str(tmp_path / "nope.scene")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open scene file" in r.stderr
    r = subprocess.run([BIN, "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown option" in r.stderr
