"""ad-hoc GPU timing of the render path used during development (not a test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spcbpt_loader
pkg = spcbpt_loader.load()
from spcbpt_optix7_b200.renderer import Renderer
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 512)
K = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
sc = pkg.scenes.scaled(pkg.scenes.cornell_scene(), 0.01)
r = Renderer(sc, w, h, K=K)
t = time.time()
st = r.preprocessing(target_samples=400000, target_Q_samples=300000, tree_samples=100000)
print("preprocessing s", time.time() - t, st)
for _ in range(3): r.render_frame()
r.ctx.synchronize()
ev = [torch.cuda.Event(True) for _ in range(4)]
n = 10
tl = te = tp = 0.0
for _ in range(n):
    ev[0].record(); r.launch_light_trace(); ev[1].record()
    r.P["sampler"] = r.ctx.lvc_process(r.lvc, r.valid, r.n_lvc)[0]; ev[2].record()
    r.launch_subframe(); ev[3].record(); r.subframe += 1
    torch.cuda.synchronize()
    tl += ev[0].elapsed_time(ev[1]); tp += ev[1].elapsed_time(ev[2]); te += ev[2].elapsed_time(ev[3])
print("per frame ms: light trace %.3f  lvc_process %.3f  eye pass %.3f  -> %.2f Msamples/s" % (tl / n, tp / n, te / n, w * h / ((tl + tp + te) / n) / 1e3))
print("launches", r.ctx.launch_count(), "mean", r.image().mean())

r.ctx.synchronize(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): r.render_frame()
r.ctx.synchronize(); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("sequential frame (wall) %.3f ms -> %.2f Msamples/s" % (dt * 1e3, w * h / dt / 1e6))
# pipelined loop: whole-frame wall time with the light trace under the eye pass
r.enable_pipelining()
for _ in range(3): r.render_frame()
r.ctx.synchronize(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): r.render_frame()
r.ctx.synchronize(); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print("pipelined frame %.3f ms -> %.2f Msamples/s" % (dt * 1e3, w * h / dt / 1e6))
