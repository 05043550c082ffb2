"""ctypes binding of the CPU oracle (oracle/_build/liborc*.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def build():
    """(Re)build the oracle with make; also builds oracle/_ref when /root/reference is present."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return " fma " in (line + " ")
    except OSError:
        pass
    return False


def lib():
    global _lib
    if _lib is not None:
        return _lib
    cands = (["liborc_fma.so"] if _cpu_has_fma() else []) + ["liborc.so"]
    path = None
    for c in cands:
        p = os.path.join(_HERE, "_build", c)
        if os.path.exists(p):
            path = p
            break
    if path is None:
        build()
        path = os.path.join(_HERE, "_build", cands[0])
    L = ctypes.CDLL(path)
    vp, i32, i64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32
    L.orc_tea.argtypes = [u32, u32, u32]
    L.orc_tea.restype = u32
    L.orc_rnd_stream.argtypes = [vp, i32, vp]
    L.orc_rnd_stream.restype = None
    L.orc_scene_create.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32]
    L.orc_scene_create.restype = vp
    L.orc_scene_destroy.argtypes = [vp]
    L.orc_scene_destroy.restype = None
    L.orc_scene_num_prims.argtypes = [vp]
    L.orc_trace_batch.argtypes = [vp, vp, i64, i32, vp, i32, i32]
    L.orc_trace_batch.restype = None
    L.orc_occlusion_batch.argtypes = [vp, vp, i64, vp, i32, i32]
    L.orc_occlusion_batch.restype = None
    _lib = L
    return L


def tea(rounds, v0, v1):
    return int(lib().orc_tea(rounds, v0 & 0xffffffff, v1 & 0xffffffff))


def rnd_stream(seed, n):
    st = np.array([seed], np.uint32)
    out = np.zeros(n, np.float32)
    lib().orc_rnd_stream(st.ctypes.data, n, out.ctypes.data)
    return out, int(st[0])


class Scene:
    """Oracle copy of a scene (same spc_mesh / spc_pbr / spc_light arrays as the product gets)."""

    def __init__(self, pkg, scene):
        self.pkg = pkg
        (meshes, mats, lights, textures, ntex), keep = pkg.pack_scene(scene)
        self.h = lib().orc_scene_create(meshes.ctypes.data, len(meshes), mats.ctypes.data, len(mats),
                                        lights.ctypes.data, len(lights), textures.ctypes.data, ntex)
        del keep

    def close(self):
        if self.h:
            lib().orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trace(self, rays, flags=1, brute=False, threads=1):
        rays = np.ascontiguousarray(rays, self.pkg.RAY)
        hits = np.zeros(rays.shape[0], self.pkg.HIT)
        lib().orc_trace_batch(self.h, rays.ctypes.data, rays.shape[0], flags, hits.ctypes.data, int(brute), threads)
        return hits

    def occlusion(self, rays, brute=False, threads=1):
        rays = np.ascontiguousarray(rays, self.pkg.RAY)
        vis = np.zeros(rays.shape[0], np.uint8)
        lib().orc_occlusion_batch(self.h, rays.ctypes.data, rays.shape[0], vis.ctypes.data, int(brute), threads)
        return vis


# ---- render path -----------------------------------------------------------------------------
def _bind_render(L):
    vp, i32 = ctypes.c_void_p, ctypes.c_int
    L.orc_light_trace.argtypes = [vp, vp, i32, i32, i32]
    L.orc_light_trace.restype = None
    L.orc_light_trace_c.argtypes = [vp, vp, i32, i32, i32, i32]
    L.orc_light_trace_c.restype = None
    L.orc_lvc_process.argtypes = [vp, vp, i32, i32, vp, vp, vp, vp, vp]
    L.orc_lvc_process.restype = None
    L.orc_eye_pass.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.orc_eye_pass.restype = None
    L.orc_bsdf.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.orc_bsdf.restype = None
    L.orc_classify.argtypes = [vp, vp, vp, i32, vp]
    L.orc_classify.restype = None
    L.orc_connect.argtypes = [vp, vp, i32, i32, vp, vp, i32, vp, vp]
    L.orc_connect.restype = None


def light_trace(scene, params, K, max_depth=0, threads=8, connections=3):
    L = lib(); _bind_render(L)
    L.orc_light_trace_c(scene.h, params.ctypes.data, K, connections, max_depth, threads)


def lvc_process(pkg, lvc, valid, K):
    L = lib(); _bind_render(L)
    n = lvc.shape[0]
    sub = np.zeros(K, pkg.SUBSPACE)
    cmfs = np.zeros(n, np.float32)
    jump = np.zeros(n, np.int32)
    vc, pc = ctypes.c_int(0), ctypes.c_int(0)
    L.orc_lvc_process(lvc.ctypes.data, valid.ctypes.data, n, K, sub.ctypes.data, cmfs.ctypes.data, jump.ctypes.data,
                      ctypes.byref(vc), ctypes.byref(pc))
    return sub, cmfs[:vc.value].copy(), jump[:vc.value].copy(), vc.value, pc.value


def eye_pass(scene, params, K, connections=3, max_depth=0, threads=8, want_first=False):
    L = lib(); _bind_render(L)
    n = int(params["width"][0]) * int(params["height"][0])
    fp = np.zeros(n, np.int32) if want_first else None
    fl = np.zeros(n, np.int32) if want_first else None
    L.orc_eye_pass(scene.h, params.ctypes.data, K, connections, max_depth, threads,
                   fp.ctypes.data if want_first else None, fl.ctypes.data if want_first else None)
    return fp, fl


def pt_pass(scene, params, K, threads=8):
    L = lib()
    L.orc_pt_pass.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    L.orc_pt_pass.restype = None
    L.orc_pt_pass(scene.h, params.ctypes.data, K, threads)


def bsdf(scene, material_id, color, N, V, Ldir, seed):
    L = lib(); _bind_render(L)
    N, V, Ldir = (np.ascontiguousarray(a, np.float32) for a in (N, V, Ldir))
    col = None if color is None else np.ascontiguousarray(color, np.float32)
    st = np.array([seed], np.uint32)
    e, p, s = np.zeros(3, np.float32), np.zeros(1, np.float32), np.zeros(3, np.float32)
    L.orc_bsdf(scene.h, material_id, None if col is None else col.ctypes.data, N.ctypes.data, V.ctypes.data, Ldir.ctypes.data,
               st.ctypes.data, e.ctypes.data, p.ctypes.data, s.ctypes.data)
    return e, float(p[0]), s, int(st[0])


def classify(pkg, tree, pos, nrm):
    L = lib(); _bind_render(L)
    tree = np.ascontiguousarray(tree, pkg.TREE_NODE)
    pos = np.ascontiguousarray(pos, np.float32)
    nrm = np.ascontiguousarray(nrm, np.float32)
    lab = np.zeros(pos.shape[0], np.int32)
    L.orc_classify(tree.ctypes.data, pos.ctypes.data, nrm.ctypes.data, pos.shape[0], lab.ctypes.data)
    return lab


def connect(pkg, scene, params, K, connections, eye, light):
    L = lib(); _bind_render(L)
    eye = np.ascontiguousarray(eye, pkg.VERTEX)
    light = np.ascontiguousarray(light, pkg.VERTEX)
    n = eye.shape[0]
    out = np.zeros((n, 3), np.float32)
    w = np.zeros(n, np.float32)
    L.orc_connect(scene.h, params.ctypes.data, K, connections, eye.ctypes.data, light.ctypes.data, n, out.ctypes.data, w.ctypes.data)
    return out, w


def set_jitter_rtl(v):
    """reproduce g++'s right-to-left evaluation of make_float2(rnd,rnd) (pinning vs libref_host only)"""
    lib().orc_set_jitter_rtl(int(v))


# ---- training path ---------------------------------------------------------------------------
def _bind_train(L):
    vp, i32, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    L.orc_pretrace.argtypes = [vp, vp, i32, i32, i32]
    L.orc_pretrace.restype = None
    for name, args, res in (("orc_ts_create", [], vp), ("orc_ts_destroy", [vp], None), ("orc_ts_gather", [vp, vp, i32, vp, i32], i32),
                            ("orc_ts_sizes", [vp, vp, vp], i32), ("orc_ts_read", [vp, vp, vp], None), ("orc_ts_reweight", [vp], None),
                            ("orc_ts_tree_points", [vp, i32, i32, vp, i32], i32), ("orc_ts_label", [vp, vp, vp], None),
                            ("orc_q_create", [i32], vp), ("orc_q_destroy", [vp], None), ("orc_q_add", [vp, vp, vp, i32], i32),
                            ("orc_q_zero_handle", [vp], None), ("orc_q_read", [vp, vp], None),
                            ("orc_ts_build_train_data", [vp, i32, vp, i32], vp), ("orc_td_destroy", [vp], None),
                            ("orc_td_sizes", [vp, vp, vp, vp], None), ("orc_td_read", [vp, vp, vp, vp, vp, vp, vp], None),
                            ("orc_ts_gamma_histogram", [vp, i32, vp], None),
                            ("orc_train_gamma", [vp, i32, vp, i32, i32, f32, vp, i32, vp], None), ("orc_gamma_to_cmf", [vp, i32, vp], None)):
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res


def pretrace(scene, params, K, max_depth=0, threads=8):
    L = lib(); _bind_train(L)
    L.orc_pretrace(scene.h, params.ctypes.data, K, max_depth, threads)


class TrainSet:
    """the accumulated training set (neat_paths / neat_conns of device_thrust.cu:428-429) and its MyThrustOp operations"""

    def __init__(self, pkg):
        self.pkg = pkg
        self.L = lib(); _bind_train(self.L)
        self.h = self.L.orc_ts_create()

    def __del__(self):
        try:
            if self.h:
                self.L.orc_ts_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def gather(self, paths, conns):
        paths = np.ascontiguousarray(paths, self.pkg.TRAIN_PATH)
        conns = np.ascontiguousarray(conns, self.pkg.TRAIN_CONN)
        return self.L.orc_ts_gather(self.h, paths.ctypes.data, paths.shape[0], conns.ctypes.data, conns.shape[0])

    def read(self):
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        self.L.orc_ts_sizes(self.h, ctypes.byref(a), ctypes.byref(b))
        paths = np.zeros(a.value, self.pkg.TRAIN_PATH)
        conns = np.zeros(b.value, self.pkg.TRAIN_CONN)
        self.L.orc_ts_read(self.h, paths.ctypes.data, conns.ctypes.data)
        return paths, conns

    def reweight(self):
        self.L.orc_ts_reweight(self.h)

    def tree_points(self, eye_side, max_size):
        cap = 1 << 22
        out = np.zeros(cap, self.pkg.DIVIDE_WEIGHT)
        n = self.L.orc_ts_tree_points(self.h, int(eye_side), max_size, out.ctypes.data, cap)
        assert n <= cap
        return out[:n].copy()

    def label(self, eye_tree, light_tree):
        e = np.ascontiguousarray(eye_tree, self.pkg.TREE_NODE)
        l = np.ascontiguousarray(light_tree, self.pkg.TREE_NODE)
        self.L.orc_ts_label(self.h, e.ctypes.data, l.ctypes.data)

    def build_train_data(self, n_samples, Q, K):
        Q = np.ascontiguousarray(Q, np.float32)
        td = self.L.orc_ts_build_train_data(self.h, n_samples, Q.ctypes.data, K)
        N, M, th = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_float(0)
        self.L.orc_td_sizes(td, ctypes.byref(N), ctypes.byref(M), ctypes.byref(th))
        out = dict(N=N.value, M=M.value, threshold=th.value, handle=td,
                   f_square=np.zeros(N.value, np.float32), pdf0=np.zeros(N.value, np.float32), P2N=np.zeros(N.value, np.int32),
                   peak=np.zeros(M.value, np.float32), label_E=np.zeros(M.value, np.int32), label_P=np.zeros(M.value, np.int32))
        self.L.orc_td_read(td, *(out[k].ctypes.data for k in ("f_square", "pdf0", "P2N", "peak", "label_E", "label_P")))
        return out

    def gamma_histogram(self, K):
        G = np.zeros((K, K), np.float32)
        self.L.orc_ts_gamma_histogram(self.h, K, G.ctypes.data)
        return G


class QEstimator:
    def __init__(self, K):
        self.K = K
        self.L = lib(); _bind_train(self.L)
        self.h = self.L.orc_q_create(K)

    def add(self, lvc, valid):
        return self.L.orc_q_add(self.h, lvc.ctypes.data, valid.ctypes.data, lvc.shape[0])

    def zero_handle(self):
        self.L.orc_q_zero_handle(self.h)

    def read(self):
        q = np.zeros(self.K, np.float32)
        self.L.orc_q_read(self.h, q.ctypes.data)
        return q


def train_gamma(td, K, G, batch_size=20000, epochs=1, lr=0.01):
    L = lib(); _bind_train(L)
    G = np.ascontiguousarray(G, np.float32).copy()
    loss = np.zeros(4096, np.float32)
    n = ctypes.c_int(0)
    L.orc_train_gamma(td["handle"], K, G.ctypes.data, batch_size, epochs, lr, loss.ctypes.data, loss.shape[0], ctypes.byref(n))
    return G, loss[:n.value].copy()


def gamma_to_cmf(G, K):
    L = lib(); _bind_train(L)
    G = np.ascontiguousarray(G, np.float32)
    out = np.zeros_like(G)
    L.orc_gamma_to_cmf(G.ctypes.data, K, out.ctypes.data)
    return out
