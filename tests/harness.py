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
    Q, cmf = random_q_gamma(K, seed + 1000)
    return eye_tree, light_tree, Q, cmf


def random_q_gamma(K, seed):
    """a random positive Q vector and a row-CDF Gamma matrix (seeded: regenerated, not stored, by the golden tests)"""
    rng = np.random.default_rng(seed)
    Q = rng.uniform(0.05, 2.0, K).astype(np.float32)
    G = rng.uniform(0.01, 1.0, (K, K)).astype(np.float32)
    G /= G.sum(1, keepdims=True)
    cmf = np.cumsum(G, axis=1, dtype=np.float32)
    cmf[:, -1] = 1.0
    return Q, cmf


GOLDEN_CFG = dict(w=48, h=40, num_core=16, core_padding=120, M_per_core=20, launch_frame=3, subframes=(0, 1, 2))


def golden_scene(pkg):
    """the 682-triangle Cornell fixture of tests/golden/render.npz"""
    return pkg.scenes.cornell_scene(wall_cells=6, box_cells=4)


def golden_render_setup(pkg, builder):
    """scene + trees + Q/Gamma used by tests/golden/make_golden.py (trees need the reference's builder)"""
    sc = golden_scene(pkg)
    K, K_light = 1000, 200
    rays = pkg.scenes.camera_rays(sc, 64, 64)
    # sample points for the trees: a deterministic cloud on the scene surfaces (vertex positions + face normals)
    P, N = [], []
    for m in sc.meshes:
        tri = m["positions"][m["indices"].astype(np.int64)]
        c = tri.mean(1)
        n = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
        n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
        P.append(c.astype(np.float32))
        N.append(n.astype(np.float32))
    del rays
    P, N = np.concatenate(P), np.concatenate(N)
    eye_tree, light_tree, Q, cmf = random_trees_and_gamma(pkg, P, N, K, K_light, builder)
    return sc, K, K_light, eye_tree, light_tree, Q, cmf, dict(GOLDEN_CFG)


UNDEFINED_ON_ORIGIN = ("color", "lastPosition", "RMIS_pointer_3", "last_lum", "lastNormalProjection", "lastSinglePdf",
                       "lastZoneId", "inBrdf", "lastBrdf", "isLastVertex_direction", "_pad")


def compare_lvc(pkg, a, valid_a, b, valid_b, exact=True, rtol=1e-5):
    """field-wise comparison of two LVCs on the fields the reference defines (depth-0 emitter vertices
    leave most fields as stack garbage, raygen.cu:172-195).  Returns a list of mismatch strings."""
    bad = []
    if not np.array_equal(valid_a, valid_b):
        return ["validState differs on %d slots" % int((valid_a != valid_b).sum())]
    v = valid_a.astype(bool)
    origin = a["depth"] == 0
    for name in pkg.VERTEX.names:
        if name in ("_pad", "inBrdf", "RMIS_pointer_3"):   # RMIS_pointer_3 is eye-side state: never written on light paths
            continue
        m = v & ~origin if name in UNDEFINED_ON_ORIGIN else v
        x, y = a[name][m], b[name][m]
        if exact or x.dtype.kind in "iu":
            if x.dtype.kind == "f":
                # bit-exact, except that NaNs only have to be NaNs on both sides (x86 and CUDA produce different NaN payloads)
                ne = (x.view(np.uint32) != y.view(np.uint32)) & ~(np.isnan(x) & np.isnan(y))
            else:
                ne = x != y
            if ne.any():
                bad.append("%s: %d of %d differ" % (name, int(ne.sum()), x.size))
        else:
            err = np.abs(x - y) / np.maximum(np.abs(y), 1e-20)
            if not (err <= rtol).all():
                bad.append("%s: max rel err %g" % (name, float(err.max())))
    return bad


class DeviceFrame:
    """The same buffers as HostFrame, in device memory (torch tensors are only the allocator here), with a
    MyParams whose pointers are device addresses -- what a reference host hands to the launch seam."""

    def __init__(self, pkg, scene, width, height, K=1000, num_core=1000, core_padding=800, M_per_core=100):
        import torch
        self.torch, self.pkg, self.scene, self.K = torch, pkg, scene, K
        self.w, self.h = width, height
        self.P = np.zeros(1, pkg.PARAMS)
        P = self.P
        eye, U, V, W = scene.camera_frame(width, height)
        P["width"], P["height"] = width, height
        P["eye"], P["U"], P["V"], P["W"] = eye, U, V, W
        dev = "cuda"
        self.accum = torch.zeros((width * height, 4), dtype=torch.float32, device=dev)
        self.frame = torch.zeros(width * height, dtype=torch.int32, device=dev)
        P["accum_buffer"], P["frame_buffer"] = self.accum.data_ptr(), self.frame.data_ptr()
        n = num_core * core_padding
        self.n_lvc = n
        self.lvc = torch.zeros(n * pkg.VERTEX.itemsize, dtype=torch.uint8, device=dev)
        self.valid = torch.zeros(n, dtype=torch.uint8, device=dev)
        lt = P["lt"]
        lt["num_core"], lt["core_padding"], lt["M_per_core"], lt["M"] = num_core, core_padding, M_per_core, num_core * M_per_core
        lt["ans"], lt["validState"] = self.lvc.data_ptr(), self.valid.data_ptr()
        P["subspace_info"]["subspaceNum"] = K
        self.keep = {}

    def _up(self, arr):
        return self.torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy()).cuda()

    def set_trees(self, eye_tree, light_tree):
        self.keep["eye_tree"] = self._up(np.ascontiguousarray(eye_tree, self.pkg.TREE_NODE))
        self.keep["light_tree"] = self._up(np.ascontiguousarray(light_tree, self.pkg.TREE_NODE))
        self.P["subspace_info"]["eye_tree"] = self.keep["eye_tree"].data_ptr()
        self.P["subspace_info"]["light_tree"] = self.keep["light_tree"].data_ptr()

    def set_q_gamma(self, Q, cmf_gamma):
        self.keep["Q"] = self._up(np.ascontiguousarray(Q, np.float32))
        self.keep["CMF"] = self._up(np.ascontiguousarray(cmf_gamma, np.float32))
        self.P["subspace_info"]["Q"] = self.keep["Q"].data_ptr()
        self.P["subspace_info"]["CMFGamma"] = self.keep["CMF"].data_ptr()

    def upload_lvc(self, lvc, valid):
        self.lvc.copy_(self._up(lvc))
        self.valid.copy_(self._up(valid))

    def lvc_host(self):
        return self.lvc.cpu().numpy().view(self.pkg.VERTEX).copy(), self.valid.cpu().numpy().copy()

    def set_sampler_record(self, rec):
        self.P["sampler"] = rec[0]

    def sampler_host(self):
        """download the SubspaceSampler arrays the record points at (through torch's raw-pointer free path: cudaMemcpy via ctypes)"""
        import ctypes
        s = self.P["sampler"][0]
        vc = int(s["vertex_count"])
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        self.torch.cuda.synchronize()
        sub = np.zeros(self.K, self.pkg.SUBSPACE)
        cmfs = np.zeros(max(vc, 1), np.float32)
        jump = np.zeros(max(vc, 1), np.int32)
        assert rt.cudaMemcpy(sub.ctypes.data, int(s["subspace"]), sub.nbytes, 2) == 0
        if vc:
            assert rt.cudaMemcpy(cmfs.ctypes.data, int(s["cmfs"]), vc * 4, 2) == 0
            assert rt.cudaMemcpy(jump.ctypes.data, int(s["jump_buffer"]), vc * 4, 2) == 0
        return sub, cmfs[:vc], jump[:vc], vc, int(s["path_count"])


def float_bits_differ(x, y):
    """boolean mask: fp32 arrays differ bitwise, NaN == NaN regardless of payload (x86 and CUDA NaNs differ)"""
    x = np.ascontiguousarray(x, np.float32)
    y = np.ascontiguousarray(y, np.float32)
    return (x.view(np.uint32) != y.view(np.uint32)) & ~(np.isnan(x) & np.isnan(y))


def setup_pretrace(frame, num_core, padding=10, iteration=1):
    """allocate the pretrace output buffers of a HostFrame (preTracer_params_setup, optixPathTracer.cpp:479-488)"""
    pkg = frame.pkg
    frame.tp = np.zeros(num_core, pkg.TRAIN_PATH)
    frame.tc = np.zeros(num_core * padding, pkg.TRAIN_CONN)
    pt = frame.P["pre_tracer"]
    pt["num_core"], pt["padding"], pt["iteration"] = num_core, padding, iteration
    pt["paths"], pt["conns"] = frame.tp.ctypes.data, frame.tc.ctypes.data


def compare_train(pkg, pa, ca, pb, cb, exact=True):
    """pretrace outputs a vs b: valid flags, then every field of the valid paths / connections"""
    bad = []
    if not np.array_equal(pa["valid"], pb["valid"]):
        return ["path valid flags differ on %d" % int((pa["valid"] != pb["valid"]).sum())]
    if not np.array_equal(ca["valid"], cb["valid"]):
        return ["conn valid flags differ on %d" % int((ca["valid"] != cb["valid"]).sum())]
    v = pa["valid"] == 1
    for k in ("contri", "sample_pdf", "fix_pdf", "begin_ind", "end_ind", "pixel_id"):
        x, y = pa[k][v], pb[k][v]
        ne = float_bits_differ(x, y) if x.dtype.kind == "f" else (x != y)
        if ne.any():
            bad.append("path.%s: %d of %d differ" % (k, int(ne.sum()), ne.size))
    v = ca["valid"] == 1
    for k in ("A_position", "B_position", "A_dir", "B_dir", "A_normal", "B_normal", "peak_pdf", "label_A", "label_B", "light_source"):
        x, y = ca[k][v], cb[k][v]
        ne = float_bits_differ(x, y) if x.dtype.kind == "f" else (x != y)
        if ne.any():
            bad.append("conn.%s: %d of %d differ" % (k, int(ne.sum()), ne.size))
    return bad


def setup_pretrace_device(df, num_core, padding=10, iteration=1):
    """device twin of setup_pretrace for a DeviceFrame"""
    torch, pkg = df.torch, df.pkg
    df.tp = torch.zeros(num_core * pkg.TRAIN_PATH.itemsize, dtype=torch.uint8, device="cuda")
    df.tc = torch.zeros(num_core * padding * pkg.TRAIN_CONN.itemsize, dtype=torch.uint8, device="cuda")
    pt = df.P["pre_tracer"]
    pt["num_core"], pt["padding"], pt["iteration"] = num_core, padding, iteration
    pt["paths"], pt["conns"] = df.tp.data_ptr(), df.tc.data_ptr()


def pretrace_host(df):
    return df.tp.cpu().numpy().view(df.pkg.TRAIN_PATH).copy(), df.tc.cpu().numpy().view(df.pkg.TRAIN_CONN).copy()


def varied_cornell(pkg):
    """Cornell fixture with a metallic box, a textured box (random 24x16 RGBA texture) and a second quad light: the material /
    texture / multi-light branches of the shading code (ColorTexSample hit_program.cu:182-198, LocalShading.h:37-53)"""
    sc = pkg.scenes.cornell_scene(wall_cells=8, box_cells=5)
    rng = np.random.default_rng(5)
    mats = pkg.scenes.make_pbr(5)
    mats[:3] = sc.materials
    mats["base_color"][3] = (0.9, 0.8, 0.3, 1); mats["metallic"][3] = 1.0; mats["roughness"][3] = 0.15
    mats["base_color"][4] = (1, 1, 1, 1); mats["roughness"][4] = 0.6
    tex = rng.integers(0, 256, (16, 24, 4), dtype=np.uint8)
    sc.textures = [tex]
    mats["base_color_tex"]["tex"][4] = 1
    sc.materials = mats
    sc.meshes[3]["material_id"] = 3     # short box: metal
    sc.meshes[4]["material_id"] = 4     # tall box: textured
    L2 = pkg.scenes.make_quad_light(1, (20.0, 300.0, 100.0), (20.0, 300.0, 200.0), (20.0, 400.0, 100.0), (6.0, 8.0, 12.0), 3, 4)
    sc.lights = np.concatenate([sc.lights, L2])
    sc.meshes.append(pkg.scenes.light_mesh(L2, 1))
    return sc
