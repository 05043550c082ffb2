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


def python_render(pkg, sc, w, h, frames, alg="SPCBPT_eye", pipelined=False):
    from spcbpt_optix7_b200.renderer import Renderer
    r = Renderer(sc, w, h, K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)
    if alg == "SPCBPT_eye":
        r.preprocessing(target_samples=40000, target_Q_samples=20000, tree_samples=20000, batch_size=20000)
        if pipelined:
            r.enable_pipelining()
        for _ in range(frames):
            r.render_frame()
    else:
        for _ in range(frames):
            r.render_frame_pt()
    return r.image().copy(), r.frame_rgba8().copy()


def test_cpp_driver_equals_python_mirror(gpu_ctx, tmp_path):
    pkg = gpu_ctx
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    path = pkg.scenes.export_scene(sc, str(tmp_path), "cb")
    w, h, frames = 96, 64, 5
    cache = tmp_path / "cb.spcscene"
    st, log = run_driver("--scene", path, "--dim=%dx%d" % (w, h), "--frames", frames, "--out", tmp_path / "a", "--save-cache", cache, "--no-pipeline", *SMALL)
    assert st["triangles"] == sc.n_triangles and st["frames"] == frames and st["kernel_launches"] > 0 and not st["pipelined"]
    img_cpp = read_pfm(tmp_path / "a.pfm")
    # the Python mirror renders the scene the C++ loader produced (material table has one entry per mesh, as in the reference)
    sc2 = pkg.scenes.load_spcscene(str(cache))
    img_py, fb_py = python_render(pkg, sc2, w, h, frames)
    assert img_cpp.shape == img_py.shape and img_py.mean() > 0.01
    assert np.array_equal(img_cpp.view(np.uint32), img_py.view(np.uint32)), "C++ driver and Python mirror disagree"
    # ... and equals the render of the original in-memory scene (export -> OBJ -> tinyobj-style load changes nothing visible)
    img_orig, _ = python_render(pkg, sc, w, h, frames)
    assert np.array_equal(img_cpp.view(np.uint32), img_orig.view(np.uint32))
    # PPM = tone-mapped frame buffer, flipped to top-down
    ppm = read_ppm(tmp_path / "a.ppm")
    assert np.array_equal(ppm, fb_py[::-1, :, :3])

    # pipelined loop (light trace of frame f+1 under the eye pass of frame f) from the cache: same image
    st2, _ = run_driver("--cache", cache, "--dim=%dx%d" % (w, h), "--frames", frames, "--out", tmp_path / "b", *SMALL)
    assert st2["pipelined"]
    assert np.array_equal(read_pfm(tmp_path / "b.pfm").view(np.uint32), img_cpp.view(np.uint32))

    # the pt comparison integrator through the same driver
    st3, _ = run_driver("--cache", cache, "--dim=%dx%d" % (w, h), "--frames", 3, "--alg", "pt", "--out", tmp_path / "c", *SMALL)
    img_pt, _ = python_render(pkg, sc2, w, h, 3, alg="pt")
    assert np.array_equal(read_pfm(tmp_path / "c.pfm").view(np.uint32), img_pt.view(np.uint32))


def test_cpp_driver_errors(gpu_ctx, tmp_path):
    r = subprocess.run([BIN, "--scene", str(tmp_path / "nope.scene")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open scene file" in r.stderr
    r = subprocess.run([BIN, "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown option" in r.stderr
