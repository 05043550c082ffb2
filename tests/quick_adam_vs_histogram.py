"""Does the Adam refinement of Gamma (train_optimal_E, device_thrust.cu:3327-3344) lower the estimator's error against the histogram Gamma
it starts from (preprocess_getGamma)?  Same training set, same trees, same Q, same render seeds; relMSE at equal spp against a pt ground
truth, de-biased by the ground truth's own noise.  A study, not a test: python tests/quick_adam_vs_histogram.py [spp]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spcbpt_loader
pkg = spcbpt_loader.load()
from spcbpt_optix7_b200.renderer import Renderer

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sc = pkg.scenes.load_spcscene(os.path.join(root, "data", "_ref", "house.spcscene"))
w, h = 960, 540


def relmse(x, ref):
    return float(np.mean((x - ref) ** 2 / (ref ** 2 + 1e-2)))


gt = Renderer(sc, w, h, K=1000)
acc = [np.zeros((h, w, 3)), np.zeros((h, w, 3))]
for c in range(16):
    gt.reset_accumulation()
    gt.ctx.set_seed_offset(7777777 + c * 256)
    for _ in range(256):
        gt.render_frame_pt()
    acc[c & 1] += np.nan_to_num(gt.image()) / 8
ref = (0.5 * (acc[0] + acc[1])).astype(np.float32)
gt_noise = float(np.mean((acc[0] - acc[1]) ** 2 / (ref.astype(np.float64) ** 2 + 1e-2))) / 4.0
out = {"image": "%dx%d" % (w, h), "spp": spp, "ground_truth": "pt 4096 spp", "ground_truth_noise_relMSE": gt_noise, "rows": []}
for name, adam in (("histogram Gamma (no Adam)", False), ("Adam-refined Gamma (100 steps)", True)):
    errs = []
    for rep in range(3):      # three independent renders each (different render seeds, same training)
        r = Renderer(sc, w, h, K=1000)
        st = r.preprocessing(adam=adam)
        r.ctx.set_seed_offset(1000 * rep)
        r.P["lt"]["launch_frame"] = 5000000 + 1000 * rep
        for _ in range(spp):
            r.render_frame()
        errs.append(relmse(np.nan_to_num(r.image()), ref) - gt_noise)
        r.ctx.close()
    out["rows"].append({"gamma": name, "relMSE_debiased": errs, "mean": float(np.mean(errs)), "loss_first": st["loss_first"], "loss_last": st["loss_last"]})
print(json.dumps(out))
