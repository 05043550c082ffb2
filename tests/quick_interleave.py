"""ad-hoc GPU experiment (not a test): several contexts rendering alternate subframes on their own streams, to see how much
of the low-occupancy tail of one frame hides under the head of the next."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spcbpt_loader
pkg = spcbpt_loader.load()
from spcbpt_optix7_b200.renderer import Renderer
cache = "data/_ref/house.spcscene"
sc = pkg.scenes.load_spcscene(cache) if os.path.exists(cache) else pkg.scenes.scaled(pkg.scenes.cornell_scene(wall_cells=72, box_cells=60), 0.01)
w, h = 1920, 1000
n_ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = 48
rs = [Renderer(sc, w, h, K=1000) for _ in range(n_ctx)]
st = rs[0].preprocessing()
si0 = rs[0].P["subspace_info"]
for r in rs[1:]:
    r.P["subspace_info"] = si0
    r.eye_tree, r.light_tree = rs[0].eye_tree, rs[0].light_tree
streams = [torch.cuda.Stream() for _ in rs]
for r, s in zip(rs, streams):
    r.ctx.set_stream(s.cuda_stream)
torch.cuda.synchronize()

def run(nf):
    lf = 100
    for f in range(nf):
        r = rs[f % n_ctx]
        lf += 1
        r.P["lt"]["launch_frame"] = lf - 1
        r.subframe = f // n_ctx
        r.render_frame()
    for r in rs:
        r.ctx.synchronize()
    torch.cuda.synchronize()

run(2 * n_ctx)
t0 = time.perf_counter()
run(frames)
dt = (time.perf_counter() - t0) / frames
print("contexts %d blocks/sm %s: %.3f ms/frame -> %.2f Msamples/s" % (n_ctx, os.environ.get("SPC_TRACE_BLOCKS_PER_SM", "dflt"), dt * 1e3, w * h / dt / 1e6))
