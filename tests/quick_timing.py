"""ad-hoc GPU timing used during development (not a test)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spcbpt_loader
pkg = spcbpt_loader.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 708
sc = pkg.scenes.heightfield_scene(n)
ctx = pkg.Context(0)
t = time.time(); ctx.upload_scene(sc); print("upload+build s", time.time() - t, ctx.bvh_stats())
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for name, rays in (("primary", pkg.scenes.camera_rays(sc, 2048, 2048)), ("random", pkg.scenes.random_rays(sc, 1 << 22, seed=1))):
    rd = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda()
    hd = torch.empty((rays.shape[0], 4), dtype=torch.float32, device="cuda")
    cnt = ctx.trace_counted(rd, rays.shape[0], hd)
    for _ in range(3): ctx.trace_device(rd, rays.shape[0], hd)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5): ctx.trace_device(rd, rays.shape[0], hd)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    nn, nt = cnt["nodes_visited"] / cnt["rays"], cnt["tris_tested"] / cnt["rays"]
    bpr = 48 + 80 * nn + 48 * nt
    print(name, "rays", rays.shape[0], "ms", ms, "Mrays/s", rays.shape[0] / ms / 1e3, "nodes/ray", nn, "tris/ray", nt, "GB/s", rays.shape[0] * bpr / ms / 1e6,
          "hit frac", float((hd[:, 3].view(torch.int32) >= 0).float().mean()))
