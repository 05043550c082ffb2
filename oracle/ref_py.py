"""ctypes binding of oracle/_ref/libref_host.so = the reference's own sources compiled for the host
(oracle/ref_shim/ref_host.cpp).  TEST INFRASTRUCTURE ONLY.  The library is built here (where
/root/reference exists) and travels to the GPU box as a prebuilt file."""
import ctypes
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_host.so")
_lib = None


def available():
    if os.path.exists(PATH):
        return True
    if os.path.isdir("/root/reference/src"):
        r = subprocess.run(["make", "-C", _HERE, "_ref/libref_host.so"], capture_output=True, text=True)
        return r.returncode == 0 and os.path.exists(PATH)
    return False


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libref_host.so is not built and /root/reference is absent")
        L = ctypes.CDLL(PATH)
        vp, i32, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32
        L.ref_layout_json.restype = ctypes.c_char_p
        L.ref_tea4.argtypes = [u32, u32]
        L.ref_tea4.restype = u32
        L.ref_tea16.argtypes = [u32, u32]
        L.ref_tea16.restype = u32
        L.ref_rnd_stream.argtypes = [vp, i32, vp]
        L.ref_rnd_stream.restype = None
        L.ref_tree_build.argtypes = [vp, i32, i32, i32, vp, i32, vp]
        L.ref_tree_index.argtypes = [vp, vp, vp, i32, vp]
        if hasattr(L, "ref_tree_load"):
            L.ref_tree_load.argtypes = [ctypes.c_char_p, vp, vp, vp, vp, i32]
        L.ref_tree_index.restype = None
        L.ref_bsdf.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
        L.ref_bsdf.restype = None
        L.ref_scene_create.argtypes = [vp, i32, vp, i32, vp, i32, vp, i32]
        L.ref_scene_destroy.restype = None
        L.ref_launch.argtypes = [vp, i32, i32, i32, i32]
        _lib = L
    return _lib


def layout():
    return json.loads(lib().ref_layout_json().decode())


def tea4(a, b):
    return int(lib().ref_tea4(a & 0xffffffff, b & 0xffffffff))


def tea16(a, b):
    return int(lib().ref_tea16(a & 0xffffffff, b & 0xffffffff))


def rnd_stream(seed, n):
    st = np.array([seed], np.uint32)
    out = np.zeros(n, np.float32)
    lib().ref_rnd_stream(st.ctypes.data, n, out.ctypes.data)
    return out, int(st[0])


def tree_build(pkg, samples, K, label_bias=0):
    """classTree::buildTreeBaseOnExistSample()(samples, K, labelBias) -> (tree_node[], max_label)"""
    samples = np.ascontiguousarray(samples, pkg.DIVIDE_WEIGHT)
    cap = 1 << 20
    out = np.zeros(cap, pkg.TREE_NODE)
    ml = ctypes.c_int(0)
    n = lib().ref_tree_build(samples.ctypes.data, samples.shape[0], K, label_bias, out.ctypes.data, cap, ctypes.byref(ml))
    assert n <= cap
    return out[:n].copy(), ml.value


def camera_uvw(eye, lookat, up, fov_y, aspect):
    """sutil::Camera(eye, lookat, up, fovY, aspect).UVWFrame -> (U, V, W), the reference's own Camera.cpp"""
    L = lib()
    L.ref_camera_uvw.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_float, ctypes.c_float, ctypes.c_void_p]
    L.ref_camera_uvw.restype = None
    a = [np.ascontiguousarray(v, np.float32) for v in (eye, lookat, up)]
    out = np.zeros(9, np.float32)
    L.ref_camera_uvw(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, float(np.float32(fov_y)), float(np.float32(aspect)), out.ctypes.data)
    return out[0:3].copy(), out[3:6].copy(), out[6:9].copy()


def tile_owner_map(w, h, num_gpus):
    """StaticWorkDistribution::getSamplePixel over all GPUs and samples -> (owner[h, w], pixels enumerated twice)"""
    L = lib()
    L.ref_tile_owner_map.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p]
    owner = np.zeros((h, w), np.int32)
    twice = L.ref_tile_owner_map(w, h, num_gpus, owner.ctypes.data)
    return owner, int(twice)


def tree_load(pkg, directory, cap=1 << 20):
    """classTree::tree_load run in `directory` (reads tree_eye.txt / tree_light.txt) -> (eye tree_node[], light tree_node[])"""
    eye, light = np.zeros(cap, pkg.TREE_NODE), np.zeros(cap, pkg.TREE_NODE)
    ne, nl = ctypes.c_int(0), ctypes.c_int(0)
    rc = lib().ref_tree_load(os.fsencode(directory), eye.ctypes.data, ctypes.byref(ne), light.ctypes.data, ctypes.byref(nl), cap)
    assert rc == 0, "ref_tree_load failed"
    return eye[:ne.value].copy(), light[:nl.value].copy()


def tree_index(pkg, tree, pos, nrm):
    tree = np.ascontiguousarray(tree, pkg.TREE_NODE)
    pos = np.ascontiguousarray(pos, np.float32)
    nrm = np.ascontiguousarray(nrm, np.float32)
    lab = np.zeros(pos.shape[0], np.int32)
    lib().ref_tree_index(tree.ctypes.data, pos.ctypes.data, nrm.ctypes.data, pos.shape[0], lab.ctypes.data)
    return lab


def bsdf(pkg, mat, N, V, L, seed):
    """Tracer::Eval / Pdf / Sample (cuProg.h:735,868,826) -> (eval3, pdf, sample3, seed_after)"""
    mat = np.ascontiguousarray(mat, pkg.PBR).reshape(1)
    N, V, L = (np.ascontiguousarray(a, np.float32) for a in (N, V, L))
    st = np.array([seed], np.uint32)
    e, p, s = np.zeros(3, np.float32), np.zeros(1, np.float32), np.zeros(3, np.float32)
    lib().ref_bsdf(mat.ctypes.data, N.ctypes.data, V.ctypes.data, L.ctypes.data, st.ctypes.data, e.ctypes.data, p.ctypes.data, s.ctypes.data)
    return e, float(p[0]), s, int(st[0])


def scene_create(pkg, scene):
    (meshes, mats, lights, textures, ntex), keep = pkg.pack_scene(scene)
    rc = lib().ref_scene_create(meshes.ctypes.data, len(meshes), mats.ctypes.data, len(mats), lights.ctypes.data, len(lights),
                                textures.ctypes.data, ntex)
    assert rc == 0
    del keep


KIND_PT, KIND_SPCBPT_EYE, KIND_LIGHT_TRACE, KIND_PRETRACE = 0, 1, 2, 3


def launch(params, kind, w, h, threads=8):
    """optixLaunch of the reference's own raygen program `kind` with a host-pointer MyParams"""
    rc = lib().ref_launch(params.ctypes.data, kind, w, h, threads)
    assert rc == 0
