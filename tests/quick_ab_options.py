"""A/B of context options on the shipped scene, in ONE process and interleaved (the GPU boxes are shared VMs: separate runs drift by
tens of per cent).  Not a test.  python tests/quick_ab_options.py [--lanes 4] [--fast] [--frames 48] [--reps 7] cfg ...
where cfg = name:opt=val,opt=val  (e.g. base: sort:sort_hits=1 notail:tail_threshold=-1)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spcbpt_loader
pkg = spcbpt_loader.load()
from spcbpt_optix7_b200.renderer import LaneRenderer

ap = argparse.ArgumentParser()
ap.add_argument("--lanes", type=int, default=4)
ap.add_argument("--fast", action="store_true")
ap.add_argument("--frames", type=int, default=48)
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--dim", default="1920x1080")
ap.add_argument("cfgs", nargs="+")
a = ap.parse_args()
w, h = (int(x) for x in a.dim.split("x"))
cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "_ref", "house.spcscene")
sc = pkg.scenes.load_spcscene(cache)
lr = LaneRenderer(sc, w, h, lanes=a.lanes, K=1000, fast=a.fast)
lr.preprocessing()
cfgs = []
for c in a.cfgs:
    name, _, rest = c.partition(":")
    cfgs.append((name, [kv.split("=") for kv in rest.split(",") if kv]))
ALL = ("sort_hits", "tail_threshold", "light_trace_mode")
times = {n: [] for n, _ in cfgs}
means = {}
for rep in range(a.reps + 1):
    for name, opts in cfgs:
        for lane in lr.lanes:
            for k in ALL:
                lane.ctx.set_option(k, 0)
            for k, v in opts:
                lane.ctx.set_option(k, int(v))
        lr.render(a.lanes)          # settle
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lr.render(a.frames)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / a.frames * 1e3
        if rep > 0:
            times[name].append(dt)
        means[name] = float(lr.image().mean())
for name, _ in cfgs:
    t = sorted(times[name])
    print(json.dumps({"cfg": name, "lanes": a.lanes, "fast": a.fast, "ms_per_frame_median": t[len(t) // 2], "min": t[0], "max": t[-1], "image_mean": means[name]}))
