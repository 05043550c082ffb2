"""spcbpt-optix7_b200 -- host-side Python mirror of the C ABI in include/spcbpt_b200.h.

The product is libspcbpt_b200.so (hand-written sm_100a CUDA + C ABI); this module only binds it
with ctypes and describes the POD structs as numpy dtypes.  There is no Python/CPU compute path:
importing works anywhere (so CPU-only tests can check symbols and layouts), but creating a
Context without a CUDA device raises.

The directory name has a hyphen (contract of the build), so import it through
``spcbpt_loader.load()`` at the repo root, which registers it as ``spcbpt_optix7_b200``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SPCBPT_LIB: development only (A/B of two builds, tests/quick_ab_*.sh); the product path is the in-tree build
LIB_PATH = os.environ.get("SPCBPT_LIB") or os.path.join(_HERE, "libspcbpt_b200.so")
# the fast-arithmetic flavour of the same library (build.py NVCC_FLAGS_FAST, csrc/shade.cuh SPC_FAST_MATH): Context(fast=True)
LIB_PATH_FAST = os.path.join(_HERE, "libspcbpt_b200_fast.so")

# --------------------------------------------------------------------------------------------
# numpy dtypes of the POD structs (byte-identical to include/spcbpt_b200.h)
# --------------------------------------------------------------------------------------------
RAY = np.dtype([("ox", "f4"), ("oy", "f4"), ("oz", "f4"), ("tmin", "f4"),
                ("dx", "f4"), ("dy", "f4"), ("dz", "f4"), ("tmax", "f4")])
HIT = np.dtype([("t", "f4"), ("u", "f4"), ("v", "f4"), ("prim", "i4")])
TEXREF = np.dtype([("texcoord", "i4"), ("_pad0", "i4"), ("tex", "u8"), ("texcoord_offset", "f4", 2),
                   ("texcoord_rotation", "f4", 2), ("texcoord_scale", "f4", 2)])
PBR = np.dtype([("base_color", "f4", 4), ("metallic", "f4"), ("roughness", "f4"), ("specular", "f4"),
                ("specularTint", "f4"), ("subsurface", "f4"), ("anisotropic", "f4"), ("sheen", "f4"),
                ("sheenTint", "f4"), ("clearcoat", "f4"), ("clearcoatGloss", "f4"),
                ("base_color_tex", TEXREF), ("metallic_roughness_tex", TEXREF), ("brdf", "u1"), ("_pad1", "u1", 7)])
LIGHT = np.dtype([("type", "i4"), ("id", "i4"), ("divLevel", "i4"), ("ssBase", "i4"), ("corner", "f4", 3),
                  ("u", "f4", 3), ("v", "f4", 3), ("emission", "f4", 3), ("normal", "f4", 3), ("area", "f4")])
VERTEX = np.dtype([("position", "f4", 3), ("normal", "f4", 3), ("flux", "f4", 3), ("color", "f4", 3),
                   ("lastPosition", "f4", 3), ("RMIS_pointer_3", "f4", 3), ("uv", "f4", 2), ("RMIS_pointer", "f4"),
                   ("last_lum", "f4"), ("lastNormalProjection", "f4"), ("pdf", "f4"), ("singlePdf", "f4"),
                   ("lastSinglePdf", "f4"), ("materialId", "i2"), ("subspaceId", "i2"), ("depth", "i2"),
                   ("lastZoneId", "i2"), ("type", "i2"), ("isOrigin", "u1"), ("inBrdf", "u1"), ("lastBrdf", "u1"),
                   ("isBrdf", "u1"), ("isLastVertex_direction", "u1"), ("_pad", "u1")])
TREE_NODE = np.dtype([("mid", "f4", 3), ("child", "i4", 8), ("label", "i4"), ("type", "i4"), ("leaf", "u1"), ("_pad", "u1", 3)])
DIVIDE_WEIGHT = np.dtype([("position", "f4", 3), ("dir", "f4", 3), ("normal", "f4", 3), ("weight", "f4")])
TRAIN_PATH = np.dtype([("contri", "f4", 3), ("sample_pdf", "f4"), ("fix_pdf", "f4"), ("begin_ind", "i4"), ("end_ind", "i4"),
                       ("choice_id", "i4"), ("pixel_id", "i4", 2), ("valid", "u1"), ("_pad", "u1", 7)])
TRAIN_CONN = np.dtype([("A_position", "f4", 3), ("B_position", "f4", 3), ("A_dir", "f4", 3), ("B_dir", "f4", 3),
                       ("A_normal", "f4", 3), ("B_normal", "f4", 3), ("peak_pdf", "f4"), ("path_id", "i4"), ("label_A", "i4"),
                       ("label_B", "i4"), ("valid", "u1"), ("light_source", "u1"), ("_pad", "u1", 2)])
SUBSPACE = np.dtype([("jump_bias", "i4"), ("id", "i4"), ("size", "i4"), ("sum_pmf", "f4"), ("Q", "f4")])
MESH = np.dtype([("positions", "u8"), ("indices", "u8"), ("texcoords", "u8"), ("n_vertices", "u4"),
                 ("n_triangles", "u4"), ("material_id", "i4"), ("light_id", "i4")])
TEXTURE = np.dtype([("rgba", "u8"), ("width", "i4"), ("height", "i4")])
EYE_STATS = np.dtype([("bounces", "i4"), ("timed", "i4"), ("closest_rays", "u8"), ("shadow_slots", "u8"), ("shadow_rays", "u8"),
                      ("visible_connections", "u8"), ("stage_ms", "f4", 8)])
STAGE_NAMES = ("trace", "shade", "sample", "shadow", "connect", "gather", "other", "total")
BVH_STATS = np.dtype([("n_triangles", "u4"), ("n_nodes", "u4"), ("n_bvh2_nodes", "u4"), ("max_depth", "u4"),
                      ("sah_cost", "f4"), ("build_ms", "f4"), ("bytes_nodes", "u8"), ("bytes_triangles", "u8")])
TRACE_COUNTERS = np.dtype([("rays", "u8"), ("nodes_visited", "u8"), ("tris_tested", "u8")])
BUFFER_VIEW = np.dtype([("data", "u8"), ("count", "u4"), ("byte_stride", "u2"), ("elmt_byte_size", "u2")])
LT_PARAMS = np.dtype([("num_core", "i4"), ("core_padding", "i4"), ("M", "i4"), ("M_per_core", "i4"), ("ans", "u8"),
                      ("validState", "u8"), ("launch_frame", "i4"), ("_pad", "i4")])
PRETRACE_PARAMS = np.dtype([("num_core", "i4"), ("padding", "i4"), ("iteration", "i4"), ("_pad", "i4"),
                            ("paths", "u8"), ("conns", "u8")])
SAMPLER = np.dtype([("LVC", "u8"), ("subspace", "u8"), ("cmfs", "u8"), ("jump_buffer", "u8"),
                    ("vertex_count", "i4"), ("path_count", "i4")])
SUBSPACE_INFO = np.dtype([("subspaceNum", "i4"), ("_pad", "i4"), ("eye_tree", "u8"), ("light_tree", "u8"),
                          ("Q", "u8"), ("CMFGamma", "u8")])
ENV_INFO = np.dtype([("tex", "u8"), ("cmf", "u8"), ("r", "f4"), ("center", "f4", 3), ("size", "i4"), ("width", "i4"),
                     ("height", "i4"), ("divLevel", "i4"), ("ssBase", "i4"), ("valid", "u1"), ("_pad", "u1", 3)])
PARAMS = np.dtype([("width", "u4"), ("height", "u4"), ("subframe_index", "u4"), ("_pad0", "u4"),
                   ("accum_buffer", "u8"), ("frame_buffer", "u8"), ("max_depth", "i4"), ("eye", "f4", 3),
                   ("U", "f4", 3), ("V", "f4", 3), ("W", "f4", 3), ("_pad1", "u4"), ("lights", BUFFER_VIEW),
                   ("materials", BUFFER_VIEW), ("miss_color", "f4", 3), ("_pad2", "u4"), ("handle", "u8"),
                   ("lt", LT_PARAMS), ("sampler", SAMPLER), ("pre_tracer", PRETRACE_PARAMS),
                   ("subspace_info", SUBSPACE_INFO), ("sky", ENV_INFO)])

EXPECTED_SIZES = {"RAY": 32, "HIT": 16, "TEXREF": 40, "PBR": 144, "LIGHT": 80, "VERTEX": 120, "TREE_NODE": 56,
                  "DIVIDE_WEIGHT": 40, "SUBSPACE": 20, "EYE_STATS": 72, "TRAIN_PATH": 48, "TRAIN_CONN": 92, "MESH": 40, "TEXTURE": 16, "BUFFER_VIEW": 16,
                  "LT_PARAMS": 40, "PRETRACE_PARAMS": 32, "SAMPLER": 40, "SUBSPACE_INFO": 40, "ENV_INFO": 56,
                  "PARAMS": 352}

LAUNCH_PT, LAUNCH_SPCBPT_EYE, LAUNCH_LIGHT_TRACE, LAUNCH_PRETRACE = 0, 1, 2, 3
RAYFLAG_NONE = 0
RAYFLAG_CULL_BACK_FACING = 1
LIGHT_QUAD = 2
VTYPE_QUAD = 1
VTYPE_HIT_LIGHT_SOURCE = 4
VTYPE_NORMALHIT = 6


class SpcError(RuntimeError):
    pass


_lib = None
_lib_fast = None


def declared_symbols():
    """Every SPC_API function declared in include/spcbpt_b200.h (parsed from the header)."""
    import re
    hdr = os.path.join(_HERE, "..", "include", "spcbpt_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"SPC_API\s+[\w\s\*]+?\b(spc_\w+)\s*\(", txt)))


def lib(fast=False):
    """Load libspcbpt_b200.so (built in-tree by build.py), or its fast-arithmetic flavour; fail loudly if it is missing."""
    global _lib, _lib_fast
    if fast:
        if _lib_fast is None:
            _lib_fast = _load(LIB_PATH_FAST)
        return _lib_fast
    if _lib is None:
        _lib = _load(LIB_PATH)
    return _lib


def _load(path):
    if not os.path.exists(path):
        raise SpcError("%s is not built: run `python spcbpt-optix7_b200/build.py` "
                       "(there is no Python or CPU fallback)" % os.path.basename(path))
    # One sequential pass over the file before mapping it: on a freshly provisioned box the library's 10 MB of device code are paged
    # in on demand, kernel by kernel, at the first launch of each (CUDA loads modules lazily) -- seconds of scattered page faults that
    # would land inside whatever phase happens to launch a kernel first (measured: 0.1 -> 1.7 s on the first tree build).
    with open(path, "rb") as fh:
        while fh.read(1 << 24):
            pass
    L = ctypes.CDLL(path)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.spc_last_error.restype = ctypes.c_char_p
    L.spc_version.restype = ctypes.c_char_p
    L.spc_create.argtypes = [i32, i32, i32, i32, ctypes.POINTER(vp)]
    L.spc_destroy.argtypes = [vp]
    L.spc_destroy.restype = None
    L.spc_set_stream.argtypes = [vp, vp]
    L.spc_synchronize.argtypes = [vp]
    L.spc_scene_upload.argtypes = [vp, vp, i32, vp, i32, vp, i32, vp, i32]
    L.spc_bvh_stats_get.argtypes = [vp, vp]
    L.spc_trace_batch.argtypes = [vp, vp, i64, i32, vp]
    L.spc_trace_batch_device.argtypes = [vp, vp, i64, i32, vp]
    L.spc_occlusion_batch.argtypes = [vp, vp, i64, vp]
    L.spc_occlusion_batch_device.argtypes = [vp, vp, i64, vp]
    L.spc_trace_batch_counted.argtypes = [vp, vp, i64, i32, vp, vp]
    L.spc_occlusion_batch_counted.argtypes = [vp, vp, i64, vp, vp]
    L.spc_launch_count.argtypes = [vp]
    L.spc_launch_count.restype = i64
    _bind_optional(L)
    return L


def _bind_optional(L):
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    table = {
        "spc_gen_camera_rays": [vp, vp, i32, i32, i32, vp],
        "spc_gen_bench_rays": [vp, i32, vp, vp, i64, vp, vp],
        "spc_set_params": [vp, vp],
        "spc_launch": [vp, i32, i32, i32],
        "spc_launch_named": [vp, ctypes.c_char_p, i32, i32],
        "spc_set_debug_outputs": [vp, vp, vp],
        "spc_eye_stats_get": [vp, vp],
        "spc_comm_unique_id": [vp],
        "spc_comm_init": [vp, i32, i32, vp],
        "spc_comm_destroy": [vp],
        "spc_comm_info": [vp, vp, vp],
        "spc_comm_barrier": [vp],
        "spc_comm_allreduce_host": [vp, vp, i32, i32, i32],
        "spc_comm_bcast_host": [vp, vp, ctypes.c_size_t, i32],
        "spc_allreduce_training_stats": [vp],
        "spc_reduce_accum": [vp, vp, i32, f32, i32],
        "spc_train_set_write": [vp, vp, i32, vp, i32],
        "spc_train_Q_write": [vp, vp, i32],
        "spc_build_tree_gpu": [vp, vp, i32, i32, i32, vp, i32, vp, vp],
        "spc_build_tree_from_training_set": [vp, i32, i32, i32, i32, vp, vp, vp, i32],
        "spc_set_option": [vp, ctypes.c_char_p, i64],
        "spc_get_option": [vp, ctypes.c_char_p, vp],
        "spc_set_seed_offset": [vp, ctypes.c_uint32],
        "spc_set_seed_mapping": [vp, ctypes.c_uint32, ctypes.c_uint32],
        "spc_set_trace_blocks": [vp, i32],
        "spc_set_tile_partition": [vp, i32, i32],
        "spc_scene_share": [vp, vp],
        "spc_merge_accum": [vp, vp, vp, i32, i32, vp, vp],
        "spc_lvc_process": [vp, vp, vp, i32, vp],
        "spc_build_tree": [vp, i32, i32, i32, vp, i32, vp],
        "spc_valid_sample_gather": [vp, vp, i32, vp, i32, vp],
        "spc_sample_reweight": [vp],
        "spc_get_tree_points": [vp, i32, i32, vp, i32, vp],
        "spc_tree_to_device": [vp, i32, vp, i32, vp],
        "spc_preprocess_getQ": [vp, vp, vp, i32, i32, vp, vp],
        "spc_Q_zero_handle": [vp],
        "spc_node_label": [vp, vp, vp],
        "spc_build_optimal_E_train_data": [vp, i32],
        "spc_preprocess_getGamma": [vp, vp],
        "spc_train_optimal_E": [vp, i32, i32, f32, vp, vp, i32, vp],
        "spc_Gamma2CMFGamma": [vp, vp, vp],
        "spc_train_set_size": [vp, vp, vp],
        "spc_train_set_read": [vp, vp, vp],
        "spc_train_data_read": [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
        "spc_train_reset": [vp],
        "spc_device_alloc": [vp, ctypes.c_size_t, vp],
        "spc_device_free": [vp, vp],
        "spc_upload": [vp, vp, vp, ctypes.c_size_t],
        "spc_download": [vp, vp, vp, ctypes.c_size_t],
    }
    for name, args in table.items():
        if hasattr(L, name):
            getattr(L, name).argtypes = args


def _ptr(a):
    """Address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


def pack_scene(scene):
    """Turn a scenes.SceneData into the (meshes, materials, lights, textures) struct arrays of the
    C ABI.  Returns (arrays, keepalive): keepalive must outlive the call that consumes the arrays."""
    keep = []
    meshes = np.zeros(len(scene.meshes), MESH)
    for i, m in enumerate(scene.meshes):
        pos = np.ascontiguousarray(m["positions"], np.float32)
        idx = np.ascontiguousarray(m["indices"], np.uint32)
        uv = None if m.get("texcoords") is None else np.ascontiguousarray(m["texcoords"], np.float32)
        keep += [pos, idx, uv]
        meshes[i]["positions"] = pos.ctypes.data
        meshes[i]["indices"] = idx.ctypes.data
        meshes[i]["texcoords"] = 0 if uv is None else uv.ctypes.data
        meshes[i]["n_vertices"] = pos.shape[0]
        meshes[i]["n_triangles"] = idx.shape[0]
        meshes[i]["material_id"] = m.get("material_id", 0)
        meshes[i]["light_id"] = m.get("light_id", -1)
    textures = np.zeros(max(len(scene.textures), 1), TEXTURE)
    for i, t in enumerate(scene.textures):
        px = np.ascontiguousarray(t, np.uint8)
        keep.append(px)
        textures[i]["rgba"] = px.ctypes.data
        textures[i]["height"], textures[i]["width"] = px.shape[0], px.shape[1]
    mats = np.ascontiguousarray(scene.materials)
    lights = np.ascontiguousarray(scene.lights)
    keep += [meshes, textures, mats, lights]
    return (meshes, mats, lights, textures, len(scene.textures)), keep


def build_tree(samples, K, label_bias=0):
    """classTree::buildTreeBaseOnExistSample()(samples, K, labelBias) on the host -> (tree_node[], max_label)"""
    L = lib()
    samples = np.ascontiguousarray(samples, DIVIDE_WEIGHT)
    cap = 1 << 16
    while True:
        out = np.zeros(cap, TREE_NODE)
        ml = ctypes.c_int(0)
        n = L.spc_build_tree(samples.ctypes.data, samples.shape[0], K, label_bias, out.ctypes.data, cap, ctypes.byref(ml))
        if n < 0:
            raise SpcError("spc_build_tree failed (%d): %s" % (n, L.spc_last_error().decode()))
        if n <= cap:
            return out[:n].copy(), ml.value
        cap = n


class Context:
    """RAII wrapper of spc_context.  Mirrors the calls a reference host makes at its two seams
    (sutil::Scene launch seam and MyThrustOp post-processing seam)."""

    def __init__(self, device=0, K=0, K_light=0, connections=0, fast=False):
        self._L = lib(fast)
        self.fast = bool(fast)
        h = ctypes.c_void_p()
        rc = self._L.spc_create(device, K, K_light, connections, ctypes.byref(h))
        if rc != 0:
            raise SpcError("spc_create failed (%d): %s" % (rc, self._L.spc_last_error().decode()))
        self.h = h
        self.device = device
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self._L.spc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise SpcError("%s failed (%d): %s" % (what, rc, self._L.spc_last_error().decode()))

    def call(self, name, *args):
        """Generic checked call: ctx.call('spc_xxx', ...) -> raises SpcError on non-zero status."""
        fn = getattr(self._L, name)
        self._ck(fn(self.h, *[_ptr(a) if not isinstance(a, (int, float)) else a for a in args]), name)

    # -- plumbing -------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._ck(self._L.spc_set_stream(self.h, ctypes.c_void_p(cuda_stream)), "spc_set_stream")

    def synchronize(self):
        self._ck(self._L.spc_synchronize(self.h), "spc_synchronize")

    def launch_count(self):
        return int(self._L.spc_launch_count(self.h))

    # -- scene ----------------------------------------------------------------------------
    def upload_scene(self, scene):
        (meshes, mats, lights, textures, ntex), keep = pack_scene(scene)
        self._ck(self._L.spc_scene_upload(self.h, meshes.ctypes.data, len(meshes), mats.ctypes.data, len(mats),
                                          lights.ctypes.data, len(lights), textures.ctypes.data, ntex),
                 "spc_scene_upload")
        del keep

    def share_scene(self, owner):
        """use `owner`'s uploaded scene and BVH (same device) instead of a copy: spc_scene_share"""
        self._ck(self._L.spc_scene_share(self.h, owner.h), "spc_scene_share")
        self._scene_owner = owner      # keeps the owner alive as long as this context

    def bvh_stats(self):
        s = np.zeros(1, BVH_STATS)
        self._ck(self._L.spc_bvh_stats_get(self.h, s.ctypes.data), "spc_bvh_stats_get")
        return {k: s[0][k].item() for k in BVH_STATS.names}

    # -- launch seam (sutil::Scene::switchRaygen + optixLaunch) and MyThrustOp seam -----------
    def set_params(self, params):
        """params: numpy array of dtype PARAMS (1 element) whose pointers are device addresses"""
        assert params.dtype == PARAMS and params.size == 1
        self._ck(self._L.spc_set_params(self.h, params.ctypes.data), "spc_set_params")

    def launch(self, kind, width, height):
        if isinstance(kind, str):
            self._ck(self._L.spc_launch_named(self.h, kind.encode(), width, height), "spc_launch_named")
        else:
            self._ck(self._L.spc_launch(self.h, kind, width, height), "spc_launch")

    def set_seed_offset(self, offset):
        self._ck(self._L.spc_set_seed_offset(self.h, offset), "spc_set_seed_offset")

    def set_seed_mapping(self, offset, stride):
        self._ck(self._L.spc_set_seed_mapping(self.h, offset, stride), "spc_set_seed_mapping")

    def set_tile_partition(self, gpu_idx, num_gpus):
        self._ck(self._L.spc_set_tile_partition(self.h, gpu_idx, num_gpus), "spc_set_tile_partition")

    def set_trace_blocks(self, blocks_per_sm):
        self._ck(self._L.spc_set_trace_blocks(self.h, blocks_per_sm), "spc_set_trace_blocks")

    def merge_accum(self, accum_devs, weights, n_pixels, out_accum_dev, out_frame_dev=None):
        """out = sum_k weights[k] * accum_devs[k] (+ the tone-mapped frame buffer): the read-out of a sample-partitioned render"""
        ptrs = np.array([int(_ptr(a)) for a in accum_devs], np.uint64)
        w = np.ascontiguousarray(weights, np.float32)
        self._ck(self._L.spc_merge_accum(self.h, ptrs.ctypes.data, w.ctypes.data, len(ptrs), n_pixels, _ptr(out_accum_dev), _ptr(out_frame_dev)), "spc_merge_accum")

    def set_option(self, name, value):
        self._ck(self._L.spc_set_option(self.h, name.encode(), int(value)), "spc_set_option(%s)" % name)

    def get_option(self, name):
        v = ctypes.c_int64(0)
        self._ck(self._L.spc_get_option(self.h, name.encode(), ctypes.byref(v)), "spc_get_option(%s)" % name)
        return v.value

    # -- multi-GPU: NCCL inside the library (csrc/comm.cu) ---------------------------------------
    def comm_unique_id(self):
        buf = ctypes.create_string_buffer(128)
        rc = self._L.spc_comm_unique_id(buf)
        if rc != 0:
            raise SpcError("spc_comm_unique_id failed (%d): %s" % (rc, self._L.spc_last_error().decode()))
        return buf.raw

    def comm_init(self, rank, world, id_bytes):
        self._ck(self._L.spc_comm_init(self.h, rank, world, ctypes.c_char_p(id_bytes) if id_bytes else None), "spc_comm_init")

    def comm_info(self):
        r, w = ctypes.c_int(0), ctypes.c_int(1)
        self._ck(self._L.spc_comm_info(self.h, ctypes.byref(r), ctypes.byref(w)), "spc_comm_info")
        return r.value, w.value

    def comm_barrier(self):
        self._ck(self._L.spc_comm_barrier(self.h), "spc_comm_barrier")

    def comm_allreduce_host(self, arr, op="sum"):
        """small numpy array (int32 / float32 / float64) all-reduced over the ranks, returned"""
        a = np.ascontiguousarray(arr).copy()
        dt = {"int32": 0, "float32": 1, "float64": 2}[a.dtype.name]
        self._ck(self._L.spc_comm_allreduce_host(self.h, a.ctypes.data, a.size, dt, {"sum": 0, "min": 1, "max": 2}[op]), "spc_comm_allreduce_host")
        return a

    def comm_bcast_array(self, arr, dtype, root=0):
        """numpy struct array from `root` to every rank (size first, then the bytes)"""
        rank, world = self.comm_info()
        n = np.array([arr.shape[0] if rank == root else 0], np.int32)
        self._ck(self._L.spc_comm_bcast_host(self.h, n.ctypes.data, 4, root), "spc_comm_bcast_host")
        out = np.ascontiguousarray(arr, dtype) if rank == root else np.zeros(int(n[0]), dtype)
        self._ck(self._L.spc_comm_bcast_host(self.h, out.ctypes.data, out.nbytes, root), "spc_comm_bcast_host")
        return out

    def allreduce_training_stats(self):
        self._ck(self._L.spc_allreduce_training_stats(self.h), "spc_allreduce_training_stats")

    def reduce_accum(self, accum_dev, n_pixels, weight, root=0):
        self._ck(self._L.spc_reduce_accum(self.h, _ptr(accum_dev), n_pixels, weight, root), "spc_reduce_accum")

    def train_set_write(self, paths, conns):
        paths, conns = np.ascontiguousarray(paths, TRAIN_PATH), np.ascontiguousarray(conns, TRAIN_CONN)
        self._ck(self._L.spc_train_set_write(self.h, paths.ctypes.data, paths.shape[0], conns.ctypes.data, conns.shape[0]), "spc_train_set_write")

    def train_Q_write(self, Q, acc_paths):
        Q = np.ascontiguousarray(Q, np.float32)
        self._ck(self._L.spc_train_Q_write(self.h, Q.ctypes.data, acc_paths), "spc_train_Q_write")

    def eye_stats(self):
        """work counters (+ per-stage device ms under option stage_timing) of the last eye pass"""
        s = np.zeros(1, EYE_STATS)
        self._ck(self._L.spc_eye_stats_get(self.h, s.ctypes.data), "spc_eye_stats_get")
        out = {k: int(s[0][k]) for k in EYE_STATS.names if k != "stage_ms"}
        out["stage_ms"] = {n: float(v) for n, v in zip(STAGE_NAMES, s[0]["stage_ms"])}
        return out

    def set_debug_outputs(self, first_prim_dev, first_label_dev):
        self._ck(self._L.spc_set_debug_outputs(self.h, _ptr(first_prim_dev), _ptr(first_label_dev)), "spc_set_debug_outputs")

    def lvc_process(self, lvc_dev, valid_dev, n):
        """MyThrustOp::LVC_Process -> numpy SAMPLER record (device pointers owned by the context)"""
        s = np.zeros(1, SAMPLER)
        self._ck(self._L.spc_lvc_process(self.h, _ptr(lvc_dev), _ptr(valid_dev), n, s.ctypes.data), "spc_lvc_process")
        return s

    # -- subspace training (MyThrustOp seam, part 2) ------------------------------------------
    def download(self, dev_ptr, dtype, count):
        out = np.zeros(count, dtype)
        self._ck(self._L.spc_download(self.h, out.ctypes.data, ctypes.c_void_p(int(dev_ptr)), out.nbytes), "spc_download")
        return out

    def valid_sample_gather(self, paths_dev, n_paths, conns_dev, n_conns):
        n = ctypes.c_int(0)
        self._ck(self._L.spc_valid_sample_gather(self.h, _ptr(paths_dev), n_paths, _ptr(conns_dev), n_conns, ctypes.byref(n)), "spc_valid_sample_gather")
        return n.value

    def sample_reweight(self):
        self._ck(self._L.spc_sample_reweight(self.h), "spc_sample_reweight")

    def get_tree_points(self, eye_side, max_size):
        n = ctypes.c_int(0)
        self._ck(self._L.spc_get_tree_points(self.h, int(eye_side), max_size, None, 0, ctypes.byref(n)), "spc_get_tree_points")
        out = np.zeros(max(n.value, 1), DIVIDE_WEIGHT)
        self._ck(self._L.spc_get_tree_points(self.h, int(eye_side), max_size, out.ctypes.data, out.shape[0], ctypes.byref(n)), "spc_get_tree_points")
        return out[:n.value]

    def build_tree_gpu(self, samples, K, label_bias=0):
        """spc_build_tree on the GPU (csrc/tree_build.cu): host samples in, (tree_node[], max_label) out -- the same tree, bit for bit"""
        samples = np.ascontiguousarray(samples, DIVIDE_WEIGHT)
        cap = 1 << 16
        while True:
            out = np.zeros(cap, TREE_NODE)
            ml, nn = ctypes.c_int(0), ctypes.c_int(0)
            self._ck(self._L.spc_build_tree_gpu(self.h, samples.ctypes.data, samples.shape[0], K, label_bias, out.ctypes.data, cap, ctypes.byref(ml), ctypes.byref(nn)),
                     "spc_build_tree_gpu")
            if nn.value <= cap:
                return out[:nn.value].copy(), ml.value
            cap = nn.value

    def build_tree_from_training_set(self, eye_side, max_size, subspaces, label_bias=0):
        """tree points -> tree -> installed as the context's eye / light tree, all on the device; returns (device pointer, host copy)"""
        p, nn = ctypes.c_void_p(), ctypes.c_int(0)
        self._ck(self._L.spc_build_tree_from_training_set(self.h, int(eye_side), max_size, subspaces, label_bias, ctypes.byref(p), ctypes.byref(nn), None, 0),
                 "spc_build_tree_from_training_set")
        return p.value, self.download(p.value, TREE_NODE, nn.value)

    def tree_to_device(self, eye_side, nodes):
        nodes = np.ascontiguousarray(nodes, TREE_NODE)
        p = ctypes.c_void_p()
        self._ck(self._L.spc_tree_to_device(self.h, int(eye_side), nodes.ctypes.data, nodes.shape[0], ctypes.byref(p)), "spc_tree_to_device")
        return p.value

    def preprocess_getQ(self, lvc_dev, valid_dev, n, reset=False):
        q, acc = ctypes.c_void_p(), ctypes.c_int(0)
        self._ck(self._L.spc_preprocess_getQ(self.h, _ptr(lvc_dev), _ptr(valid_dev), n, int(reset), ctypes.byref(q), ctypes.byref(acc)), "spc_preprocess_getQ")
        return q.value, acc.value

    def Q_zero_handle(self):
        self._ck(self._L.spc_Q_zero_handle(self.h), "spc_Q_zero_handle")

    def node_label(self, eye_tree_dev, light_tree_dev):
        self._ck(self._L.spc_node_label(self.h, ctypes.c_void_p(eye_tree_dev), ctypes.c_void_p(light_tree_dev)), "spc_node_label")

    def build_optimal_E_train_data(self, n_samples):
        self._ck(self._L.spc_build_optimal_E_train_data(self.h, n_samples), "spc_build_optimal_E_train_data")

    def preprocess_getGamma(self):
        g = ctypes.c_void_p()
        self._ck(self._L.spc_preprocess_getGamma(self.h, ctypes.byref(g)), "spc_preprocess_getGamma")
        return g.value

    def train_optimal_E(self, batch_size=0, epochs=0, lr=0.0):
        g, nb = ctypes.c_void_p(), ctypes.c_int(0)
        loss = np.zeros(8192, np.float32)
        self._ck(self._L.spc_train_optimal_E(self.h, batch_size, epochs, lr, ctypes.byref(g), loss.ctypes.data, loss.shape[0], ctypes.byref(nb)), "spc_train_optimal_E")
        return g.value, loss[:nb.value].copy()

    def Gamma2CMFGamma(self, gamma_dev):
        p = ctypes.c_void_p()
        self._ck(self._L.spc_Gamma2CMFGamma(self.h, ctypes.c_void_p(gamma_dev), ctypes.byref(p)), "spc_Gamma2CMFGamma")
        return p.value

    def train_set_read(self):
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        self._ck(self._L.spc_train_set_size(self.h, ctypes.byref(a), ctypes.byref(b)), "spc_train_set_size")
        paths, conns = np.zeros(a.value, TRAIN_PATH), np.zeros(b.value, TRAIN_CONN)
        self._ck(self._L.spc_train_set_read(self.h, paths.ctypes.data, conns.ctypes.data), "spc_train_set_read")
        return paths, conns

    def train_data_read(self):
        N, M, th = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_float(0)
        self._ck(self._L.spc_train_data_read(self.h, ctypes.byref(N), ctypes.byref(M), ctypes.byref(th), None, None, None, None, None, None), "spc_train_data_read")
        out = dict(N=N.value, M=M.value, threshold=th.value,
                   f_square=np.zeros(N.value, np.float32), pdf0=np.zeros(N.value, np.float32), P2N=np.zeros(N.value, np.int32),
                   peak=np.zeros(M.value, np.float32), label_E=np.zeros(M.value, np.int32), label_P=np.zeros(M.value, np.int32))
        self._ck(self._L.spc_train_data_read(self.h, None, None, None, *(out[k].ctypes.data for k in ("f_square", "pdf0", "P2N", "peak", "label_E", "label_P"))), "spc_train_data_read")
        return out

    def train_reset(self):
        self._ck(self._L.spc_train_reset(self.h), "spc_train_reset")

    # -- ray batches (host buffers: the e2e path) --------------------------------------------
    def trace(self, rays, flags=RAYFLAG_CULL_BACK_FACING):
        rays = np.ascontiguousarray(rays, RAY)
        hits = np.zeros(rays.shape[0], HIT)
        self._ck(self._L.spc_trace_batch(self.h, rays.ctypes.data, rays.shape[0], flags, hits.ctypes.data), "spc_trace_batch")
        return hits

    def occlusion(self, rays):
        rays = np.ascontiguousarray(rays, RAY)
        vis = np.zeros(rays.shape[0], np.uint8)
        self._ck(self._L.spc_occlusion_batch(self.h, rays.ctypes.data, rays.shape[0], vis.ctypes.data), "spc_occlusion_batch")
        return vis

    # -- ray batches (device pointers: torch tensors or raw addresses) -----------------------
    def trace_device(self, rays_dev, n, hits_dev, flags=RAYFLAG_CULL_BACK_FACING):
        self._ck(self._L.spc_trace_batch_device(self.h, _ptr(rays_dev), n, flags, _ptr(hits_dev)), "spc_trace_batch_device")

    def occlusion_device(self, rays_dev, n, vis_dev):
        self._ck(self._L.spc_occlusion_batch_device(self.h, _ptr(rays_dev), n, _ptr(vis_dev)), "spc_occlusion_batch_device")

    def trace_counted(self, rays_dev, n, hits_dev, flags=RAYFLAG_CULL_BACK_FACING):
        c = np.zeros(1, TRACE_COUNTERS)
        self._ck(self._L.spc_trace_batch_counted(self.h, _ptr(rays_dev), n, flags, _ptr(hits_dev), c.ctypes.data), "spc_trace_batch_counted")
        return {k: int(c[0][k]) for k in TRACE_COUNTERS.names}

    def occlusion_counted(self, rays_dev, n, vis_dev):
        c = np.zeros(1, TRACE_COUNTERS)
        self._ck(self._L.spc_occlusion_batch_counted(self.h, _ptr(rays_dev), n, _ptr(vis_dev), c.ctypes.data), "spc_occlusion_batch_counted")
        return {k: int(c[0][k]) for k in TRACE_COUNTERS.names}


from . import scenes  # noqa: E402,F401
