"""The reference fixes its training / launch schedule by literals scattered over optixPathTracer.cpp, optixPathTracer.h and
device_thrust.cu.  This test reads those literals from the reference tree (authoring container only) and
compares them with the defaults of the three hosts of this repository: the C++ driver (host/spcbpt_main.cpp Options), the Python
twin (renderer.py) and the library (spc_create / spc_train_optimal_E defaults)."""
import inspect
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/OptiXPathTracer"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree (authoring container)")


def _int(pattern, text):
    m = re.search(pattern, text)
    assert m, pattern
    return int(m.group(1))


def test_schedule_literals_equal_the_reference(pkg):
    main = open(os.path.join(REF, "optixPathTracer.cpp")).read()
    hdr = open(os.path.join(REF, "optixPathTracer.h")).read()
    thrust = open(os.path.join(REF, "cuda_thrust", "device_thrust.cu")).read()
    train_fn = thrust[thrust.index("void train_optimal_E(thrust::device_ptr<float>& E_ptr)"):][:1500]
    ref = dict(
        lt_per_core=_int(r"lt_params\.M_per_core\s*=\s*(\d+);", main), lt_padding=_int(r"lt_params\.core_padding\s*=\s*(\d+);", main),
        lt_cores=_int(r"lt_params\.num_core\s*=\s*(\d+);", main), pre_cores=_int(r"pr_params\.num_core\s*=\s*(\d+);", main),
        pre_padding=_int(r"pr_params\.padding\s*=\s*(\d+);", main), train_samples=_int(r"const int target_sample_count\s*=\s*(\d+);", main),
        q_samples=_int(r"const int target_Q_samples\s*=\s*(\d+);", main),
        tree_samples=_int(r"get_weighted_point_for_tree_building\(true,\s*(\d+)\)", main),
        K=_int(r"#define NUM_SUBSPACE (\d+)", hdr), connections=_int(r"#define CONNECTION_N (\d+)", hdr),
        batch=_int(r"theta\.fit\((\d+),\s*epoches", train_fn), epochs=_int(r"int epoches\s*=\s*(\d+);", train_fn),
        width=_int(r"int32_t\s+width\s*=\s*(\d+);", main), height=_int(r"int32_t\s+height\s*=\s*(\d+);", main))
    m = re.search(r"#define NUM_SUBSPACE_LIGHTSOURCE \(int\(([0-9.]+) \* NUM_SUBSPACE\)\)", hdr)
    k_light_frac = float(m.group(1))
    lr = float(re.search(r"float lr\s*=\s*([0-9.]+);", train_fn).group(1))
    assert ref == dict(lt_per_core=100, lt_padding=800, lt_cores=1000, pre_cores=10000, pre_padding=10, train_samples=2000000, q_samples=2000000,
                       tree_samples=100000, K=1000, connections=3, batch=20000, epochs=1, width=1920, height=1000) and k_light_frac == 0.2 and lr == 0.01

    # C++ driver: Options defaults
    cpp = open(os.path.join(ROOT, "host", "spcbpt_main.cpp")).read()
    opts = cpp[cpp.index("struct Options {"):cpp.index("};", cpp.index("struct Options {"))]
    for key in ("lt_per_core", "lt_padding", "lt_cores", "pre_cores", "pre_padding", "train_samples", "q_samples", "tree_samples", "K", "connections", "batch", "epochs", "width", "height"):
        assert _int(r"\b%s\s*=\s*(\d+)" % key, opts) == ref[key], key
    assert float(re.search(r"float lr\s*=\s*([0-9.]+)f;", opts).group(1)) == lr
    assert re.search(r"K_light\s*=\s*0\b", opts)                                    # 0 -> the library's default below

    # Python twin: Renderer.__init__ and Renderer.preprocessing defaults
    from spcbpt_optix7_b200.renderer import Renderer
    init = {k: v.default for k, v in inspect.signature(Renderer.__init__).parameters.items()}
    pre = {k: v.default for k, v in inspect.signature(Renderer.preprocessing).parameters.items()}
    assert (init["lt_num_core"], init["lt_core_padding"], init["lt_M_per_core"], init["pretrace_num_core"], init["pretrace_padding"]) == \
        (ref["lt_cores"], ref["lt_padding"], ref["lt_per_core"], ref["pre_cores"], ref["pre_padding"])
    assert (pre["target_samples"], pre["target_Q_samples"], pre["tree_samples"], pre["batch_size"], pre["epochs"], pre["lr"]) == \
        (ref["train_samples"], ref["q_samples"], ref["tree_samples"], ref["batch"], ref["epochs"], lr)

    # library: spc_create maps 0 to the reference's compile-time constants, spc_train_optimal_E maps 0 to the reference's literals
    api = open(os.path.join(ROOT, "spcbpt-optix7_b200", "csrc", "api.cu")).read()
    assert _int(r"if \(K == 0\) K = (\d+);", api) == ref["K"] and _int(r"if \(connections == 0\) connections = (\d+);", api) == ref["connections"]
    assert float(re.search(r"if \(K_light == 0\) K_light = int\(([0-9.]+) \* K\);", api).group(1)) == k_light_frac
    api_train = open(os.path.join(ROOT, "spcbpt-optix7_b200", "csrc", "api_train.cu")).read()
    m = re.search(r"batch_size > 0 \? batch_size : (\d+), epochs > 0 \? epochs : (\d+), lr > 0 \? lr : ([0-9.]+)f", api_train)
    assert (int(m.group(1)), int(m.group(2)), float(m.group(3))) == (ref["batch"], ref["epochs"], lr)
