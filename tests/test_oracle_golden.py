"""CPU tests: the oracle (oracle/*.cpp, the CPU restatement of the reference's path) against the
golden vectors in tests/golden/, which were produced by the reference's OWN sources compiled for
the host (tests/golden/make_golden.py -> oracle/_ref/libref_host.so).  Everything is bit-exact:
both sides are scalar fp32 C++ with the same libm and no contraction."""
import hashlib
import json
import os

import numpy as np
import pytest

from harness import GOLDEN_CFG, HostFrame, compare_lvc, compare_train, golden_scene, random_q_gamma, setup_pretrace

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


def test_pretrace_vs_reference_golden(pkg, orc, golden_frame):
    """oracle __raygen__TrainData restatement == the reference's own program (tests/golden/train.npz)"""
    g, sc, osc, fr, K = golden_frame
    t = np.load(os.path.join(GOLD, "train.npz"))
    setup_pretrace(fr, 3000, 10, iteration=5)
    orc.set_jitter_rtl(1)
    try:
        orc.pretrace(osc, fr.P, K, threads=4)
    finally:
        orc.set_jitter_rtl(0)
    bad = compare_train(pkg, fr.tp, fr.tc, t["paths"], t["conns"])
    assert not bad, bad
    assert t["paths"]["valid"].sum() > 1000


def test_training_oracle_properties(pkg, orc, golden_frame):
    """structure of the MyThrustOp restatement on the golden training paths: gather keeps order and fixes indices, the Gamma
    histogram rows are normalised, the CDF is monotone and ends at 1, the trainer lowers its own loss"""
    g, sc, osc, fr, K = golden_frame
    t = np.load(os.path.join(GOLD, "train.npz"))
    ts = orc.TrainSet(pkg)
    n = ts.gather(t["paths"], t["conns"])
    assert n == int(t["paths"]["valid"].sum())
    p, c = ts.read()
    assert (p["end_ind"] > p["begin_ind"]).all() and p["begin_ind"][0] == 0 and (p["begin_ind"][1:] == p["end_ind"][:-1]).all()
    assert c.shape[0] == p["end_ind"][-1] and (c["path_id"] == np.repeat(np.arange(n), p["end_ind"] - p["begin_ind"])).all()
    ts.reweight()
    ts.label(g["eye_tree"], g["light_tree"])
    p, c = ts.read()
    assert c["label_A"].max() < K and c["label_B"].max() < K
    Q = np.full(K, 0.5, np.float32)
    td = ts.build_train_data(n, Q, K)
    G = ts.gamma_histogram(K)
    assert np.allclose(G.sum(1), 1, atol=1e-4)
    E, loss = orc.train_gamma(td, K, G, 500, 3, 0.01)
    # (this 556-unit fixture underflows sample_pdf on long paths, so a few losses are inf: only the bookkeeping is checked
    #  here; the trainer's arithmetic is checked against numpy in test_trainer_matches_numpy and on the GPU chain)
    assert loss.shape[0] == 3 * (n // 500)
    ok = np.isfinite(E).all(1)
    assert ok.mean() > 0.95 and np.allclose(E[ok].sum(1), 1, atol=1e-4)
    C = orc.gamma_to_cmf(E, K)
    assert (C[:, -1] == 1).all() and (np.diff(C[ok], axis=1) >= -1e-7).all()


def test_trainer_matches_numpy(pkg, orc):
    """train_optimal_E restatement (sigmoid/row-normalised E, 1/pdf loss, the reference's gradient formulas, Adam) against an
    independent float64 numpy implementation on a synthetic training set"""
    K, N, B = 8, 1200, 400
    rng = np.random.default_rng(0)
    paths = np.zeros(N, pkg.TRAIN_PATH)
    conns = np.zeros(N * 10, pkg.TRAIN_CONN)
    lens = rng.integers(1, 4, N)
    for i in range(N):
        paths[i]["valid"], paths[i]["begin_ind"], paths[i]["end_ind"] = 1, i * 10, i * 10 + lens[i]
        paths[i]["contri"], paths[i]["sample_pdf"], paths[i]["fix_pdf"] = rng.uniform(0.1, 1, 3), rng.uniform(0.5, 2), rng.uniform(0.1, 0.5)
        for k in range(lens[i]):
            c = conns[i * 10 + k]
            c["valid"], c["peak_pdf"], c["label_A"], c["label_B"] = 1, rng.uniform(0.1, 3), rng.integers(0, K), rng.integers(0, K)
    ts = orc.TrainSet(pkg)
    ts.gather(paths, conns)
    td = ts.build_train_data(N, np.ones(K, np.float32), K)
    G0 = ts.gamma_histogram(K)
    E, loss = orc.train_gamma(td, K, G0, B, 1, 0.01)

    def sig(x):
        return 1 / (1 + np.exp(-x))
    theta = -np.log(1 / G0.astype(np.float64) - 1)
    m, v = np.zeros_like(theta), np.zeros_like(theta)
    f2, pdf0, peak = (td[k].astype(np.float64) for k in ("f_square", "pdf0", "peak"))
    P2N, lE, lP = td["P2N"], td["label_E"], td["label_P"]
    ref_loss = []
    for b in range(N // B):
        bs, bn = b * B, P2N[b * B]
        seg = (td["M"] if bs + B >= N else P2N[bs + B]) - bn
        S = sig(theta)
        Es = S.sum(1, keepdims=True)
        Em = S / Es * 0.8 + 0.2 / K
        pdf = np.zeros(B)
        np.add.at(pdf, lP[bn:bn + seg] - bs, peak[bn:bn + seg] * Em.reshape(-1)[lE[bn:bn + seg]])
        pdf += pdf0[bs:bs + B]
        ref_loss.append((f2[bs:bs + B] / pdf).mean())
        d = -f2[bs:bs + B] / pdf ** 2
        dE = np.zeros(K * K)
        np.add.at(dE, lE[bn:bn + seg], peak[bn:bn + seg] * d[lP[bn:bn + seg] - bs])
        dE = dE.reshape(K, K)
        dEsum = (-(Em * Es) / Es / Es * dE).sum(1, keepdims=True)
        g = S * (1 - S) * dEsum + ((Em * Es) * (1 - (Em * Es)) / Es) * dE
        t = b + 1
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        theta -= 0.01 * (m / (1 - 0.9 ** t)) / (np.sqrt(v / (1 - 0.999 ** t)) + 1e-8)
    assert np.allclose(loss, ref_loss, rtol=1e-5)
    S = sig(theta)
    assert np.abs(S / S.sum(1, keepdims=True) - E).max() < 1e-5
