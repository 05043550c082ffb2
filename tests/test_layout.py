"""CPU tests: struct layouts, exported symbols, loader behaviour (no compute calls)."""
import ctypes
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_numpy_dtypes_match_header_sizes(pkg):
    for name, size in pkg.EXPECTED_SIZES.items():
        assert getattr(pkg, name).itemsize == size, name


def test_params_offsets(pkg):
    # field offsets of MyParams probed from the reference headers (SURVEY.md section 8 intro)
    want = dict(width=0, height=4, subframe_index=8, accum_buffer=16, frame_buffer=24, max_depth=32, eye=36,
                U=48, V=60, W=72, lights=88, materials=104, miss_color=120, handle=136, lt=144, sampler=184,
                pre_tracer=224, subspace_info=256, sky=296)
    for k, off in want.items():
        assert pkg.PARAMS.fields[k][1] == off, k


def test_vertex_offsets(pkg):
    want = dict(position=0, normal=12, flux=24, color=36, lastPosition=48, RMIS_pointer_3=60, uv=72,
                RMIS_pointer=80, last_lum=84, lastNormalProjection=88, pdf=92, singlePdf=96, lastSinglePdf=100,
                materialId=104, subspaceId=106, depth=108, lastZoneId=110, type=112, isOrigin=114, inBrdf=115,
                lastBrdf=116, isBrdf=117, isLastVertex_direction=118)
    for k, off in want.items():
        assert pkg.VERTEX.fields[k][1] == off, k


def test_layout_matches_reference_headers(pkg):
    """golden/ref_layout.json is produced by oracle/_ref/libref_host.so = sizeof/offsetof evaluated on
    the reference's own headers (tests/golden/make_golden.py)."""
    ref = json.load(open(os.path.join(GOLD, "ref_layout.json")))
    assert ref["sizeof"]["BDPTVertex"] == pkg.VERTEX.itemsize
    assert ref["sizeof"]["MyParams"] == pkg.PARAMS.itemsize
    assert ref["sizeof"]["Light"] == pkg.LIGHT.itemsize
    assert ref["sizeof"]["MaterialData::Pbr"] == pkg.PBR.itemsize
    assert ref["sizeof"]["tree_node"] == pkg.TREE_NODE.itemsize
    assert ref["sizeof"]["Subspace"] == pkg.SUBSPACE.itemsize
    assert ref["sizeof"]["divide_weight"] == pkg.DIVIDE_WEIGHT.itemsize
    for k, off in ref["offsetof"]["MyParams"].items():
        assert pkg.PARAMS.fields[k][1] == off, k
    for k, off in ref["offsetof"]["BDPTVertex"].items():
        assert pkg.VERTEX.fields[k][1] == off, k
    for k, off in ref["offsetof"]["Light"].items():
        assert pkg.LIGHT.fields[k][1] == off, k
    for k, off in ref["offsetof"]["Pbr"].items():
        assert pkg.PBR.fields[k][1] == off, k


@pytest.mark.parametrize("flavour", ["exact", "fast"])
def test_library_exports_every_declared_symbol(pkg, flavour):
    path = pkg.LIB_PATH if flavour == "exact" else pkg.LIB_PATH_FAST
    assert os.path.exists(path), "build the extension first (python spcbpt-optix7_b200/build.py)"
    L = ctypes.CDLL(path)
    syms = pkg.declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product must refuse to run (no oracle, no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.SpcError) as e:
        pkg.Context(0)
    assert "no CUDA device" in str(e.value) or "NO_DEVICE" in str(e.value) or "-3" in str(e.value)


def test_product_does_not_reference_oracle():
    """the product (package + C ABI sources) must not include, link, load or import anything under oracle/"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = re.compile(r'liborc|libref_host|orc_py|ref_py|load_oracle|#\s*include\s*[<"][^">]*(oracle|orc_)[^">]*[">]|dlopen\((?!"libnccl)|import\s+oracle|from\s+oracle')
    for sub in ("spcbpt-optix7_b200", "include", "host"):
        for dp, _, files in os.walk(os.path.join(root, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    m = bad.search(txt)
                    assert not m, "%s: %s" % (f, m.group(0))


def test_tree_build_matches_reference_golden(pkg):
    """spc_build_tree (host code, as in the reference) against the tree the reference's own
    classTree::buildTreeBaseOnExistSample produced on the same 2000 samples (tests/golden/tree.npz)"""
    g = np.load(os.path.join(GOLD, "tree.npz"))
    tree, max_label = pkg.build_tree(g["samples"], 16, 0)
    ref = g["tree"]
    assert tree.shape[0] == ref.shape[0] and max_label == int(g["max_label"])
    assert np.array_equal(tree["leaf"], ref["leaf"]) and np.array_equal(tree["label"], ref["label"])
    inner = ref["leaf"] == 0     # mid / child / type of leaves are uninitialised memory in the reference
    assert np.array_equal(tree["type"][inner], ref["type"][inner])
    assert np.array_equal(tree["child"][inner], ref["child"][inner])
    assert np.array_equal(tree["mid"][inner].view(np.uint32), ref["mid"][inner].view(np.uint32))
