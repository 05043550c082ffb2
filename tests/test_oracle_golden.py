"""CPU tests: the oracle (oracle/*.cpp, the CPU restatement of the reference's path) against the
golden vectors in tests/golden/, which were produced by the reference's OWN sources compiled for
the host (tests/golden/make_golden.py -> oracle/_ref/libref_host.so).  Everything is bit-exact:
both sides are scalar fp32 C++ with the same libm and no contraction."""
import hashlib
import json
import os

import numpy as np
import pytest

from harness import GOLDEN_CFG, HostFrame, compare_lvc, golden_scene, random_q_gamma

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_rng_known_answers(orc):
    g = json.load(open(os.path.join(GOLD, "rng.json")))
    for e in g["tea4"]:
        s = orc.tea(4, e["v0"], e["v1"])
        assert s == e["seed"]
        draws, state = orc.rnd_stream(s, len(e["rnd"]))
        assert [float(x) for x in draws] == e["rnd"]
        assert state == e["state_after"]
    for e in g["tea16"]:
        assert orc.tea(16, e["v0"], e["v1"]) == e["seed"]


def test_rng_survey_vectors(orc):
    # vectors probed from the reference's random.h during the survey
    assert orc.tea(4, 0, 0) == 1576399551
    assert orc.tea(4, 5, 7) == 2032901574
    assert orc.tea(16, 5, 7) == 1769051054
    d, st = orc.rnd_stream(1576399551, 3)
    assert np.allclose(d, [0.294449925, 0.695515215, 0.897309542], rtol=0, atol=1e-9) and st == 3873814036


def test_classification_labels(pkg, orc):
    g = np.load(os.path.join(GOLD, "tree.npz"))
    lab = orc.classify(pkg, g["tree"], g["probe_p"], g["probe_n"])
    assert np.array_equal(lab, g["labels"])
    assert len(np.unique(lab)) > 8


def test_bsdf(pkg, orc):
    g = np.load(os.path.join(GOLD, "bsdf.npz"))
    sc = pkg.scenes.cornell_scene(wall_cells=1, box_cells=1)
    sc.materials = g["mat"]
    for m in sc.meshes:
        if m["light_id"] < 0:
            m["material_id"] = 0
    osc = orc.Scene(pkg, sc)
    for i in range(g["mat"].shape[0]):
        e, p, s, sa = orc.bsdf(osc, i, None, g["N"][i], g["V"][i], g["L"][i], int(g["seed"][i]))
        assert np.array_equal(e.view(np.uint32), g["eval"][i].view(np.uint32)), i
        assert np.float32(p).view(np.uint32) == g["pdf"][i].view(np.uint32), i
        assert np.array_equal(s.view(np.uint32), g["sample"][i].view(np.uint32)), i
        assert sa == int(g["seed_after"][i])


@pytest.fixture(scope="module")
def golden_frame(pkg, orc):
    g = np.load(os.path.join(GOLD, "render.npz"))
    sc = golden_scene(pkg)
    K = 1000
    Q, cmf = random_q_gamma(K, 1000)
    assert hashlib.sha256(Q.tobytes()).hexdigest() == str(g["q_sha"]) and hashlib.sha256(cmf.tobytes()).hexdigest() == str(g["cmf_sha"])
    c = GOLDEN_CFG
    fr = HostFrame(pkg, sc, c["w"], c["h"], K=K, num_core=c["num_core"], core_padding=c["core_padding"], M_per_core=c["M_per_core"])
    fr.set_trees(g["eye_tree"], g["light_tree"])
    fr.set_q_gamma(Q, cmf)
    fr.P["lt"]["launch_frame"] = c["launch_frame"]
    return g, sc, orc.Scene(pkg, sc), fr, K


def test_light_trace_lvc(pkg, orc, golden_frame):
    g, sc, osc, fr, K = golden_frame
    orc.light_trace(osc, fr.P, K, threads=4)
    bad = compare_lvc(pkg, fr.lvc, fr.valid, g["lvc"], g["valid"], exact=True)
    assert not bad, bad
    assert int(fr.valid.sum()) == int(g["vc"])


def test_lvc_process_properties(pkg, orc, golden_frame):
    g, sc, osc, fr, K = golden_frame
    sub, cmfs, jump, vc, pc = orc.lvc_process(pkg, g["lvc"], g["valid"], K)
    assert vc == int(g["vc"]) and pc == int(g["pc"])
    assert np.array_equal(jump, g["jump"]) and np.array_equal(cmfs.view(np.uint32), g["cmfs"].view(np.uint32))
    # structural properties of MyThrustOp::LVC_Process (device_thrust.cu:241-332)
    assert sub["size"].sum() == vc and (np.cumsum(sub["size"]) - sub["size"] == sub["jump_bias"]).all()
    ids = g["lvc"]["subspaceId"][jump]
    for s in np.nonzero(sub["size"])[0][:50]:
        b, n = sub["jump_bias"][s], sub["size"][s]
        assert (ids[b:b + n] == s).all() and (np.diff(jump[b:b + n]) > 0).all()
        assert (np.diff(cmfs[b:b + n]) >= 0).all() and abs(cmfs[b + n - 1] - 1) < 1e-6
    assert pc == int(((g["lvc"]["depth"] == 0) & (g["valid"] == 1)).sum())


def test_eye_pass_accum(pkg, orc, golden_frame):
    g, sc, osc, fr, K = golden_frame
    fr.lvc[:] = g["lvc"]
    fr.valid[:] = g["valid"]
    fr.set_sampler(g["sub"].copy(), g["cmfs"].copy(), g["jump"].copy(), int(g["vc"]), int(g["pc"]))
    orc.set_jitter_rtl(1)   # the golden file comes from a g++ build: make_float2(rnd,rnd) is evaluated right to left there
    try:
        for k, sf in enumerate(GOLDEN_CFG["subframes"]):
            fr.P["subframe_index"] = sf
            orc.eye_pass(osc, fr.P, K, 3, 0, threads=4)
            assert np.array_equal(fr.accum.view(np.uint32), g["accum"][k].view(np.uint32)), "subframe %d" % sf
            assert np.array_equal(fr.frame, g["frame"][k]), "frame buffer, subframe %d" % sf
    finally:
        orc.set_jitter_rtl(0)
    assert fr.accum[:, :3].mean() > 0.01
