"""ad-hoc GPU timing of the traversal kernels on the microbench ray sets (not a test).  usage: quick_timing.py [side]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spcbpt_loader
pkg = spcbpt_loader.load()
side = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sc = pkg.scenes.heightfield_scene(708)
ctx = pkg.Context(0)
ctx.upload_scene(sc)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
n = side * side
eye, U, V, W = sc.camera_frame(side, side)
cam = np.concatenate([eye, U, V, W]).astype(np.float32)
A = torch.empty((n, 8), dtype=torch.float32, device="cuda"); B = torch.empty_like(A); C = torch.empty_like(A)
hits = torch.empty((n, 4), dtype=torch.float32, device="cuda"); vis = torch.empty((n,), dtype=torch.uint8, device="cuda")
ctx.call("spc_gen_camera_rays", cam, side, side, 1, A)
ctx.trace_device(A, n, hits)
ctx.call("spc_gen_bench_rays", 1, A, hits, n, B, None)
ctx.call("spc_gen_bench_rays", 2, A, hits, n, C, None)
ctx.synchronize()
def timeit(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
tA = timeit(lambda: ctx.trace_device(A, n, hits))
tB = timeit(lambda: ctx.trace_device(B, n, hits))
tC = timeit(lambda: ctx.occlusion_device(C, n, vis))
print("env", {k: v for k, v in os.environ.items() if k.startswith("SPC_")}, "A %.2f ms %.0f Mr/s | B %.2f ms %.0f Mr/s | C %.2f ms %.0f Mr/s" % (tA, n / tA / 1e3, tB, n / tB / 1e3, tC, n / tC / 1e3))
