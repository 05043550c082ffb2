"""Equal-time / equal-spp relMSE of SPCBPT against the `pt` integrator on the shipped scene (BASELINE.json configs[2]; not a test).
Ground truth: `pt` at --gt-spp samples per pixel.  relMSE = mean((I - I*)^2 / (I*^2 + 1e-2)).
Writes one JSON line; run on a B200:  python tests/quick_equal_time.py [--width 960 --height 500 --gt-spp 16384 --seconds 2]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spcbpt_loader
pkg = spcbpt_loader.load()
from spcbpt_optix7_b200.renderer import LaneRenderer, Renderer

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=960)
ap.add_argument("--height", type=int, default=500)
ap.add_argument("--gt-spp", type=int, default=16384)
ap.add_argument("--seconds", type=float, default=2.0)
ap.add_argument("--lanes", type=int, default=3)
a = ap.parse_args()
cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "_ref", "house.spcscene")
sc = pkg.scenes.load_spcscene(cache) if os.path.exists(cache) else pkg.scenes.scaled(pkg.scenes.cornell_scene(), 0.01)
w, h = a.width, a.height


def relmse(x, ref):
    e = (x - ref) ** 2 / (ref ** 2 + 1e-2)
    return float(np.mean(e[np.isfinite(e)]))


# Ground truth from samples disjoint from both contestants, in chunks of 256 spp: the reference's pt integrator has no NaN guard
# (raygen.cu:120-137 adds payload.result unchecked), so a 0/0 sample poisons a pixel's running mean for good; poisoned chunks of
# a pixel are dropped (a handful of pixel-chunks in 10^10 samples).
gt = Renderer(sc, w, h, K=1000)
t0 = time.perf_counter()
acc = np.zeros((h, w, 3), np.float64)
cnt = np.zeros((h, w, 1), np.float64)
chunk = 256
dropped = 0
for c in range(max(1, a.gt_spp // chunk)):
    gt.reset_accumulation()
    gt.ctx.set_seed_offset(7777777 + c * chunk)
    for _ in range(chunk):
        gt.render_frame_pt()
    img = gt.image()
    ok = np.isfinite(img).all(-1, keepdims=True)
    acc += np.where(ok, img, 0.0)
    cnt += ok
    dropped += int((~ok).sum())
ref = (acc / np.maximum(cnt, 1)).astype(np.float32)
gt_s = time.perf_counter() - t0

out = {"scene": os.path.basename(cache) if os.path.exists(cache) else "cornell", "width": w, "height": h, "gt": "pt %d spp (%.1f s)" % (a.gt_spp, gt_s),
       "gt_mean": float(ref.mean()), "gt_dropped_pixel_chunks": dropped, "rows": []}
# pt, sequential loop, for `seconds`
pt = Renderer(sc, w, h, K=1000)
for _ in range(8):
    pt.render_frame_pt()
pt.reset_accumulation()
pt.ctx.synchronize()
t0 = time.perf_counter()
n = 0
while time.perf_counter() - t0 < a.seconds:
    for _ in range(16):
        pt.render_frame_pt()
    pt.ctx.synchronize()
    n += 16
dt = time.perf_counter() - t0
out["rows"].append({"alg": "pt", "spp": n, "seconds": dt, "relMSE": relmse(pt.image(), ref), "mean": float(pt.image().mean())})
# SPCBPT with frame lanes for `seconds` (training time reported separately: it is paid once per scene)
t0 = time.perf_counter()
lr = LaneRenderer(sc, w, h, lanes=a.lanes, K=1000)
st = lr.preprocessing()
torch.cuda.synchronize()
pre_s = time.perf_counter() - t0
t0 = time.perf_counter()
n = 0
while time.perf_counter() - t0 < a.seconds:
    lr.render(4 * a.lanes)
    n += 4 * a.lanes
dt = time.perf_counter() - t0
img = lr.image()
out["rows"].append({"alg": "SPCBPT_eye (%d lanes)" % a.lanes, "spp": n, "seconds": dt, "preprocess_s": pre_s, "relMSE": relmse(img, ref), "mean": float(img.mean())})
# equal spp: pt at the spp SPCBPT reached
pt2 = Renderer(sc, w, h, K=1000)
for _ in range(n):
    pt2.render_frame_pt()
out["rows"].append({"alg": "pt (equal spp)", "spp": n, "relMSE": relmse(pt2.image(), ref), "mean": float(pt2.image().mean())})
print(json.dumps(out))
