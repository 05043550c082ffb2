"""ctypes binding of oracle/_ref/libref_device.so = the reference's OWN device programs (raygen.cu, hit_program.cu, cuProg.h,
rmis.h) compiled unmodified for sm_100a with --use_fast_math against a device stub <optix.h> (oracle/ref_shim/ref_device.cu);
optixTrace is this repository's traversal.  BASELINE / TEST INFRASTRUCTURE ONLY: the GPU-side reference arm of bench.py and the
`-m gpu` tests; it needs a CUDA device.  One scene per process (file-static state, like the reference application)."""
import ctypes
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_device.so")
KIND_PT, KIND_SPCBPT_EYE, KIND_LIGHT_TRACE, KIND_PRETRACE = 0, 1, 2, 3
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(PATH)
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        L.refdev_last_error.restype = ctypes.c_char_p
        L.refdev_scene_create.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, i32]
        L.refdev_launch.argtypes = [vp, i32, i32, i32, vp]
        L.refdev_scene_destroy.restype = None
        _lib = L
    return _lib


def _ck(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, lib().refdev_last_error().decode()))


def scene_create(pkg, ctx, scene):
    """`ctx`: a product context that holds the same scene (its BVH is borrowed for optixTrace)"""
    (meshes, mats, lights, textures, ntex), keep = pkg.pack_scene(scene)
    _ck(lib().refdev_scene_create(ctx.h, meshes.ctypes.data, len(meshes), mats.ctypes.data, len(mats), lights.ctypes.data, len(lights),
                                  textures.ctypes.data, ntex), "refdev_scene_create")
    del keep


def scene_destroy():
    lib().refdev_scene_destroy()


def launch(params, kind, w, h, stream=0):
    _ck(lib().refdev_launch(params.ctypes.data, kind, w, h, ctypes.c_void_p(stream)), "refdev_launch")


class ReferenceLoop:
    """The reference application's per-frame loop (optixPathTracer.cpp:791-822) on the GPU with the reference's own code:
    launchLVCTrace = `light trace` programs + MyThrustOp::LVC_Process (oracle/_ref/libref_thrust.so, T2), launchSubframe =
    `SPCBPT_eye` programs; a device synchronisation after every launch as the reference does (CUDA_SYNC_CHECK).  The trained
    state (trees, Q, CMFGamma) is whatever `renderer` (a product Renderer on the same scene, K = 1000) holds."""

    def __init__(self, pkg, renderer, ref_thrust):
        import torch
        self.torch, self.pkg, self.r, self.rt = torch, pkg, renderer, ref_thrust
        assert renderer.K == lib().refdev_num_subspace() == ref_thrust.lib().ref_thrust_num_subspace(), "the reference compiles NUM_SUBSPACE in"
        scene_create(pkg, renderer.ctx, renderer.scene)
        self.P = renderer.P.copy()
        dev = renderer.dev
        self.accum = torch.zeros((renderer.w * renderer.h, 4), dtype=torch.float32, device=dev)
        self.frame = torch.zeros(renderer.w * renderer.h, dtype=torch.int32, device=dev)
        self.lvc = torch.zeros_like(renderer.lvc)
        self.valid = torch.zeros_like(renderer.valid)
        self.P["accum_buffer"], self.P["frame_buffer"] = self.accum.data_ptr(), self.frame.data_ptr()
        self.P["lt"]["ans"], self.P["lt"]["validState"] = self.lvc.data_ptr(), self.valid.data_ptr()
        self.P["lt"]["launch_frame"] = 500000
        self.subframe = 0
        self.stage_s = {"light_trace": 0.0, "lvc_process": 0.0, "eye": 0.0}

    def render_frame(self):
        torch, P = self.torch, self.P
        t0 = time.perf_counter()
        P["lt"]["launch_frame"] += 1
        launch(P, KIND_LIGHT_TRACE, int(P["lt"]["num_core"][0]), 1)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        s = np.zeros(1, self.pkg.SAMPLER)
        self.rt.lib().ref_thrust_lvc_process(self.lvc.data_ptr(), self.valid.data_ptr(), self.r.n_lvc, s.ctypes.data)
        P["sampler"] = s[0]
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        P["subframe_index"] = self.subframe
        launch(P, KIND_SPCBPT_EYE, self.r.w, self.r.h)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        self.subframe += 1
        self.stage_s["light_trace"] += t1 - t0
        self.stage_s["lvc_process"] += t2 - t1
        self.stage_s["eye"] += t3 - t2

    def render_frame_pt(self):
        self.P["subframe_index"] = self.subframe
        launch(self.P, KIND_PT, self.r.w, self.r.h)
        self.torch.cuda.synchronize()
        self.subframe += 1

    def image(self):
        self.torch.cuda.synchronize()
        return self.accum.cpu().numpy()[:, :3].reshape(self.r.h, self.r.w, 3)

    def close(self):
        scene_destroy()
