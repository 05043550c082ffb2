"""Host-side frame state shared by the tests: the buffers a reference host allocates around the
launch seam (optixPathTracer.cpp:260-310, 462-490) and the MyParams struct that points at them.
`HostFrame` keeps everything in numpy (for the oracle / the reference-on-host shim)."""
import numpy as np


class HostFrame:
    def __init__(self, pkg, scene, width, height, K=1000, num_core=1000, core_padding=800, M_per_core=100):
        self.pkg, self.scene, self.K = pkg, scene, K
        self.w, self.h = width, height
        self.P = np.zeros(1, pkg.PARAMS)
        P = self.P
        eye, U, V, W = scene.camera_frame(width, height)
        P["width"], P["height"] = width, height
        P["eye"], P["U"], P["V"], P["W"] = eye, U, V, W
        self.accum = np.zeros((width * height, 4), np.float32)
        self.frame = np.zeros(width * height, np.uint32)
        P["accum_buffer"] = self.accum.ctypes.data
        P["frame_buffer"] = self.frame.ctypes.data
        n = num_core * core_padding
        self.lvc = np.zeros(n, pkg.VERTEX)
        self.valid = np.zeros(n, np.uint8)
        lt = P["lt"]
        lt["num_core"], lt["core_padding"], lt["M_per_core"], lt["M"] = num_core, core_padding, M_per_core, num_core * M_per_core
        lt["ans"], lt["validState"] = self.lvc.ctypes.data, self.valid.ctypes.data
        P["subspace_info"]["subspaceNum"] = K
        self.keep = {}

    def set_trees(self, eye_tree, light_tree):
        self.keep["eye_tree"] = np.ascontiguousarray(eye_tree, self.pkg.TREE_NODE)
        self.keep["light_tree"] = np.ascontiguousarray(light_tree, self.pkg.TREE_NODE)
        self.P["subspace_info"]["eye_tree"] = self.keep["eye_tree"].ctypes.data
        self.P["subspace_info"]["light_tree"] = self.keep["light_tree"].ctypes.data

    def set_q_gamma(self, Q, cmf_gamma):
        self.keep["Q"] = np.ascontiguousarray(Q, np.float32)
        self.keep["CMF"] = np.ascontiguousarray(cmf_gamma, np.float32)
        self.P["subspace_info"]["Q"] = self.keep["Q"].ctypes.data
        self.P["subspace_info"]["CMFGamma"] = self.keep["CMF"].ctypes.data

    def set_sampler(self, sub, cmfs, jump, vertex_count, path_count):
        self.keep["sub"], self.keep["cmfs"], self.keep["jump"] = sub, cmfs, jump
        s = self.P["sampler"]
        s["LVC"] = self.lvc.ctypes.data
        s["subspace"] = sub.ctypes.data
        s["cmfs"] = cmfs.ctypes.data
        s["jump_buffer"] = jump.ctypes.data
        s["vertex_count"], s["path_count"] = vertex_count, path_count


def random_trees_and_gamma(pkg, points, normals, K, K_light, builder, seed=0):
    """classification trees over `points` + a random positive Q and row-CDF Gamma (test fixtures)."""
    rng = np.random.default_rng(seed)
    s = np.zeros(points.shape[0], pkg.DIVIDE_WEIGHT)
    s["position"], s["normal"] = points, normals
    s["weight"] = rng.uniform(0.2, 1.0, points.shape[0]).astype(np.float32)
    eye_tree, _ = builder(pkg, s, K, 0)
    s["weight"] = rng.uniform(0.2, 1.0, points.shape[0]).astype(np.float32)
    light_tree, _ = builder(pkg, s, K - K_light, 0)
    Q = rng.uniform(0.05, 2.0, K).astype(np.float32)
    G = rng.uniform(0.01, 1.0, (K, K)).astype(np.float32)
    G /= G.sum(1, keepdims=True)
    cmf = np.cumsum(G, axis=1, dtype=np.float32)
    cmf[:, -1] = 1.0
    return eye_tree, light_tree, Q, cmf
