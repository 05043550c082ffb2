"""Device-side classification-tree build (csrc/tree_build.cu: GPU nearest-centre labelling + level-synchronous octree) against
  * the tree the reference's own classTree::buildTreeBaseOnExistSample produced (tests/golden/tree.npz), and
  * the host builder (spc_build_tree, itself pinned to the reference at K = 1000 in tests/test_oracle_vs_ref.py)
on clustered, degenerate and training-set inputs: node count, leaf flags, labels, split types, children and midpoints bit-equal."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _same_tree(a, b, what):
    assert a.shape == b.shape, "%s: %d vs %d nodes" % (what, a.shape[0], b.shape[0])
    assert np.array_equal(a["leaf"], b["leaf"]), what
    assert np.array_equal(a["label"], b["label"]), "%s: %d labels differ" % (what, int((a["label"] != b["label"]).sum()))
    inner = b["leaf"] == 0          # mid / child / type of leaves are uninitialised memory in the reference
    assert np.array_equal(a["type"][inner], b["type"][inner]), what
    assert np.array_equal(a["child"][inner], b["child"][inner]), what
    assert np.array_equal(a["mid"][inner].view(np.uint32), b["mid"][inner].view(np.uint32)), what


def test_device_tree_equals_reference_golden(gpu_ctx):
    pkg = gpu_ctx
    g = np.load(os.path.join(GOLD, "tree.npz"))
    ctx = pkg.Context(0, K=64, K_light=12)
    tree, max_label = ctx.build_tree_gpu(g["samples"], 16, 0)
    assert max_label == int(g["max_label"])
    _same_tree(tree, g["tree"], "golden")
    ctx.close()


def test_device_tree_equals_host_builder(gpu_ctx):
    pkg = gpu_ctx
    ctx = pkg.Context(0)
    rng = np.random.default_rng(11)
    n = 60000
    s = np.zeros(n, pkg.DIVIDE_WEIGHT)
    c = rng.uniform(-4, 4, (40, 3))
    s["position"] = (c[rng.integers(0, 40, n)] + rng.normal(0, 0.3, (n, 3))).astype(np.float32)
    nn = rng.normal(0, 1, (n, 3))
    s["normal"] = (nn / np.linalg.norm(nn, axis=1, keepdims=True)).astype(np.float32)
    s["dir"] = s["normal"]
    s["weight"] = rng.uniform(0, 1, n).astype(np.float32) ** 3
    for K, bias in ((1000, 0), (800, 0), (52, 3), (7, 0)):
        a, ma = ctx.build_tree_gpu(s, K, bias)
        b, mb = pkg.build_tree(s, K, bias)
        assert ma == mb
        _same_tree(a, b, "clustered K=%d" % K)
    # degenerate inputs: zero weights (emitter endpoints on the light side), duplicated points, axis-aligned normals, all-negative axis
    t = s[:20000].copy()
    t["weight"][::3] = 0
    t["position"][5000:9000] = t["position"][5000]
    t["normal"][:10000] = (0, 1, 0)
    t["position"][:, 2] = -np.abs(t["position"][:, 2]) - 1
    a, ma = ctx.build_tree_gpu(t, 300, 0)
    b, mb = pkg.build_tree(t, 300, 0)
    assert ma == mb
    _same_tree(a, b, "degenerate")
    ctx.close()


def test_training_set_trees_on_device_equal_host_path(gpu_ctx):
    """the production path: points straight from the device-resident training set (spc_build_tree_from_training_set)"""
    pkg = gpu_ctx
    from spcbpt_optix7_b200.renderer import Renderer
    sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=12, box_cells=8), 0.01)
    r = Renderer(sc, 96, 64, K=64, K_light=12, lt_num_core=100, lt_core_padding=300, lt_M_per_core=40, pretrace_num_core=20000)
    n = 0
    while n < 60000:
        n += r.launch_pretrace()
    r.ctx.sample_reweight()
    for eye_side, sub in ((True, 64), (False, 52)):
        dev, tree = r.ctx.build_tree_from_training_set(eye_side, 30000, sub, 0)
        host, _ = pkg.build_tree(r.ctx.get_tree_points(eye_side, 30000), sub, 0)
        _same_tree(tree, host, "training set, eye_side=%s" % eye_side)
        assert tree.shape[0] > 100 and dev
