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
