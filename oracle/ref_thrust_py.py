"""ctypes binding of oracle/_ref/libref_thrust.so = the reference's OWN MyThrustOp library
(cuda_thrust/device_thrust.cu) compiled with nvcc for sm_100a by oracle/Makefile.  TEST INFRASTRUCTURE ONLY; it needs a
CUDA device, so only `-m gpu` tests import it.  The library is stateful (file-static thrust vectors): one training set
per process, as in the reference application."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libref_thrust.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(PATH)
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        for name, args, res in (("ref_thrust_num_subspace", [], i32), ("ref_thrust_lvc_process", [vp, vp, i32, vp], None),
                                ("ref_thrust_valid_sample_gather", [vp, i32, vp, i32], i32), ("ref_thrust_sample_reweight", [], None),
                                ("ref_thrust_get_tree_points", [i32, i32, vp, i32], i32), ("ref_thrust_tree_to_device", [i32, vp, i32], vp),
                                ("ref_thrust_get_Q", [vp, vp, i32, i32], i32), ("ref_thrust_Q_zero_handle", [], None), ("ref_thrust_Q_ptr", [], vp),
                                ("ref_thrust_node_label", [vp, vp], None), ("ref_thrust_build_train_data", [i32], None),
                                ("ref_thrust_get_gamma", [], vp), ("ref_thrust_train_gamma", [], vp), ("ref_thrust_gamma_to_cmf", [], vp),
                                ("ref_thrust_download", [vp, vp, ctypes.c_size_t], i32), ("ref_thrust_set_sizes", [vp, vp], None),
                                ("ref_thrust_set_read", [vp, vp], None), ("ref_thrust_train_data_sizes", [vp, vp], None),
                                ("ref_thrust_train_data_read", [vp, vp, vp, vp, vp, vp], None)):
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def download(ptr, dtype, count):
    out = np.zeros(count, dtype)
    assert lib().ref_thrust_download(ctypes.c_void_p(int(ptr)), out.ctypes.data, out.nbytes) == 0
    return out


def lvc_process(pkg, lvc_ptr, valid_ptr, n):
    s = np.zeros(1, pkg.SAMPLER)
    lib().ref_thrust_lvc_process(lvc_ptr, valid_ptr, n, s.ctypes.data)
    vc = int(s["vertex_count"][0])
    K = lib().ref_thrust_num_subspace()
    sub = download(s["subspace"][0], pkg.SUBSPACE, K)
    cmfs = download(s["cmfs"][0], np.float32, max(vc, 1))[:vc]
    jump = download(s["jump_buffer"][0], np.int32, max(vc, 1))[:vc]
    return sub, cmfs, jump, vc, int(s["path_count"][0])


def train_set_read(pkg):
    a, b = ctypes.c_int(0), ctypes.c_int(0)
    lib().ref_thrust_set_sizes(ctypes.byref(a), ctypes.byref(b))
    paths, conns = np.zeros(a.value, pkg.TRAIN_PATH), np.zeros(b.value, pkg.TRAIN_CONN)
    lib().ref_thrust_set_read(paths.ctypes.data, conns.ctypes.data)
    return paths, conns


def tree_points(pkg, eye_side, max_size):
    cap = 1 << 22
    out = np.zeros(cap, pkg.DIVIDE_WEIGHT)
    n = lib().ref_thrust_get_tree_points(int(eye_side), max_size, out.ctypes.data, cap)
    assert n <= cap
    return out[:n].copy()


def train_data_read():
    N, M = ctypes.c_int(0), ctypes.c_int(0)
    lib().ref_thrust_train_data_sizes(ctypes.byref(N), ctypes.byref(M))
    out = dict(N=N.value, M=M.value, f_square=np.zeros(N.value, np.float32), pdf0=np.zeros(N.value, np.float32), P2N=np.zeros(N.value, np.int32),
               peak=np.zeros(M.value, np.float32), label_E=np.zeros(M.value, np.int32), label_P=np.zeros(M.value, np.int32))
    lib().ref_thrust_train_data_read(*(out[k].ctypes.data for k in ("f_square", "pdf0", "P2N", "peak", "label_E", "label_P")))
    return out
