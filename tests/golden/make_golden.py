"""Generates tests/golden/* from the reference's OWN sources compiled for the host
(oracle/_ref/libref_host.so, built by oracle/Makefile from /root/reference).  Run here, in the
authoring container (the reference tree does not exist on the GPU box); the outputs are committed.

    python tests/golden/make_golden.py

Files:
  ref_layout.json   sizeof/offsetof/#defines evaluated on the reference headers
  rng.json          tea<4>/tea<16> known answers + LCG streams (src/cuda/random.h)
  tree.npz          classTree::buildTreeBaseOnExistSample on 2000 seeded samples + tree_index labels
  bsdf.npz          Tracer::Eval / Pdf / Sample on 256 seeded inputs
  train.npz         the reference's __raygen__TrainData (3000 launch indices, iteration 5) on the same fixture
  render.npz        the reference's raygen/closest-hit programs on a 682-triangle Cornell fixture:
                    LVC of a light-trace launch, accum buffers of three SPCBPT_eye subframes
                    (intersection = the contract of oracle/orc_scene.cpp, see ref_host.cpp header)
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import spcbpt_loader  # noqa: E402
from harness import HostFrame, golden_render_setup  # noqa: E402


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_py", os.path.join(ROOT, "oracle", "ref_py.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


def main():
    pkg = spcbpt_loader.load()
    ref = load_ref()
    assert ref.available(), "needs /root/reference (oracle/_ref/libref_host.so)"

    json.dump(ref.layout(), open(os.path.join(HERE, "ref_layout.json"), "w"), indent=1, sort_keys=True)

    # ---- RNG
    pairs = [(0, 0), (1, 0), (0, 1), (5, 7), (131328, 3), (960960, 1), (2073599, 999), (0xffffffff, 0xffffffff)]
    rng = {"tea4": [], "tea16": []}
    for a, b in pairs:
        s = ref.tea4(a, b)
        draws, state = ref.rnd_stream(s, 8)
        rng["tea4"].append({"v0": a, "v1": b, "seed": s, "rnd": [float(x) for x in draws], "state_after": state})
        rng["tea16"].append({"v0": a, "v1": b, "seed": ref.tea16(a, b)})
    json.dump(rng, open(os.path.join(HERE, "rng.json"), "w"), indent=1)

    # ---- classification tree
    g = np.random.default_rng(42)
    s = np.zeros(2000, pkg.DIVIDE_WEIGHT)
    s["position"] = g.uniform(-5, 5, (2000, 3)).astype(np.float32)
    n = g.normal(0, 1, (2000, 3))
    s["normal"] = (n / np.linalg.norm(n, axis=1, keepdims=True)).astype(np.float32)
    s["weight"] = g.uniform(0.1, 1, 2000).astype(np.float32)
    tree, max_label = ref.tree_build(pkg, s, 16, 0)
    probe_p = g.uniform(-6, 6, (4096, 3)).astype(np.float32)
    pn = g.normal(0, 1, (4096, 3))
    probe_n = (pn / np.linalg.norm(pn, axis=1, keepdims=True)).astype(np.float32)
    labels = ref.tree_index(pkg, tree, probe_p, probe_n)
    np.savez_compressed(os.path.join(HERE, "tree.npz"), samples=s, tree=tree, max_label=max_label, probe_p=probe_p,
                        probe_n=probe_n, labels=labels)

    # ---- BSDF
    m = pkg.scenes.make_pbr(256)
    m["base_color"][:, :3] = g.uniform(0.02, 1, (256, 3))
    m["metallic"] = g.choice([0.0, 0.3, 1.0], 256)
    m["roughness"] = g.choice([0.0, 0.05, 0.2, 0.5, 1.0], 256)
    m["clearcoat"] = g.choice([0.0, 0.5], 256)
    m["sheen"] = g.choice([0.0, 0.7], 256)
    m["subsurface"] = g.choice([0.0, 0.4], 256)
    m["specularTint"] = g.choice([0.0, 0.6], 256)

    def unit(k):
        v = g.normal(0, 1, (k, 3))
        return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    N, V, L = unit(256), unit(256), unit(256)
    flip = (N * V).sum(1) < 0
    V[flip] *= -1            # the programs always call with dot(N,V) >= 0 (N is flipped toward the ray)
    seeds = g.integers(0, 2 ** 32, 256, dtype=np.uint64).astype(np.uint32)
    ev, pdf, smp, seed_after = np.zeros((256, 3), np.float32), np.zeros(256, np.float32), np.zeros((256, 3), np.float32), np.zeros(256, np.uint32)
    for i in range(256):
        e, p, sdir, sa = ref.bsdf(pkg, m[i:i + 1], N[i], V[i], L[i], int(seeds[i]))
        ev[i], pdf[i], smp[i], seed_after[i] = e, p, sdir, sa
    np.savez_compressed(os.path.join(HERE, "bsdf.npz"), mat=m, N=N, V=V, L=L, seed=seeds, eval=ev, pdf=pdf, sample=smp, seed_after=seed_after)

    # ---- render stages on the small Cornell fixture (K = NUM_SUBSPACE = 1000 is compiled into the reference)
    sc, K, K_light, eye_tree, light_tree, Q, cmf, cfg = golden_render_setup(pkg, ref.tree_build)
    ref.scene_create(pkg, sc)
    fr = HostFrame(pkg, sc, cfg["w"], cfg["h"], K=K, num_core=cfg["num_core"], core_padding=cfg["core_padding"], M_per_core=cfg["M_per_core"])
    fr.set_trees(eye_tree, light_tree)
    fr.set_q_gamma(Q, cmf)
    fr.P["lt"]["launch_frame"] = cfg["launch_frame"]
    ref.launch(fr.P, ref.KIND_LIGHT_TRACE, cfg["num_core"], 1, threads=4)
    # the sampler is built by the oracle's LVC_Process restatement (the reference's own is thrust/CUDA,
    # device_thrust.cu:241-332) and stored so that the eye pass below is reproducible from the file alone
    orc = spcbpt_loader.load_oracle()
    sub, cmfs, jump, vc, pc = orc.lvc_process(pkg, fr.lvc, fr.valid, K)
    fr.set_sampler(sub, cmfs, jump, vc, pc)
    accums, frames = [], []
    for sf in cfg["subframes"]:
        fr.P["subframe_index"] = sf
        ref.launch(fr.P, ref.KIND_SPCBPT_EYE, cfg["w"], cfg["h"], threads=4)
        accums.append(fr.accum.copy())
        frames.append(fr.frame.copy())
    lvc = fr.lvc.copy()
    # zero what the reference leaves undefined (stack garbage) so the file is deterministic: invalid slots,
    # and on depth-0 emitter vertices every field init_vertex_from_lightSample (raygen.cu:172-195) does not set
    lvc[fr.valid == 0] = np.zeros(1, pkg.VERTEX)
    und = ("color", "lastPosition", "RMIS_pointer_3", "last_lum", "lastNormalProjection", "lastSinglePdf", "lastZoneId", "_pad",
           "inBrdf", "lastBrdf", "isLastVertex_direction")
    o = (fr.valid == 1) & (lvc["depth"] == 0)
    for k in und:
        lvc[k][o] = 0
    nz = (fr.valid == 1) & (lvc["depth"] > 0)
    for k in ("inBrdf", "_pad", "RMIS_pointer_3"):   # RMIS_pointer_3: eye-side only, garbage on light paths
        lvc[k][nz] = 0
    import hashlib
    # Q / CMFGamma are regenerated from their seed by the tests (harness.random_q_gamma(1000, 1000)); only a digest is stored
    np.savez_compressed(os.path.join(HERE, "render.npz"), eye_tree=eye_tree, light_tree=light_tree,
                        q_sha=hashlib.sha256(Q.tobytes()).hexdigest(), cmf_sha=hashlib.sha256(cmf.tobytes()).hexdigest(), lvc=lvc, valid=fr.valid,
                        sub=sub, cmfs=cmfs, jump=jump, vc=vc, pc=pc, accum=np.stack(accums), frame=np.stack(frames))
    # ---- training tracer: the reference's __raygen__TrainData on the same fixture (jitter drawn right-to-left: g++ build)
    from harness import setup_pretrace
    setup_pretrace(fr, 3000, 10, iteration=5)
    ref.launch(fr.P, ref.KIND_PRETRACE, 3000, 1, threads=4)
    tp, tc = fr.tp.copy(), fr.tc.copy()
    tp[tp["valid"] == 0] = np.zeros(1, pkg.TRAIN_PATH)      # undefined content of invalid records
    tc[tc["valid"] == 0] = np.zeros(1, pkg.TRAIN_CONN)
    tc["path_id"] = 0                                        # set later by valid_sample_gather
    tp["choice_id"] = 0                                      # never written by the reference
    np.savez_compressed(os.path.join(HERE, "train.npz"), paths=tp, conns=tc)
    ref.lib().ref_scene_destroy()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
