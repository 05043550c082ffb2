"""Where does the preprocessing time go, call by call?  (not a test)  python tests/quick_preprocess_phases.py [reps]"""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spcbpt_loader
pkg = spcbpt_loader.load()
from spcbpt_optix7_b200.renderer import Renderer
sc = pkg.scenes.load_spcscene(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "_ref", "house.spcscene"))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for rep in range(reps):
    r = Renderer(sc, 1920, 1080, K=1000, K_light=200)
    acc = collections.OrderedDict()
    def wrap(obj, name):
        f = getattr(obj, name)
        def g(*a, **k):
            t0 = time.perf_counter()
            out = f(*a, **k)
            obj.synchronize() if hasattr(obj, "synchronize") else None
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
            return out
        setattr(obj, name, g)
    for nm in ("sample_reweight", "build_tree_from_training_set", "preprocess_getQ", "allreduce_training_stats", "Q_zero_handle", "node_label",
               "build_optimal_E_train_data", "preprocess_getGamma", "train_optimal_E", "Gamma2CMFGamma", "valid_sample_gather", "launch", "set_params"):
        if hasattr(r.ctx, nm):
            wrap(r.ctx, nm)
    t0 = time.perf_counter()
    st = r.preprocessing()
    tot = time.perf_counter() - t0
    print("rep %d total %.3f  pretrace %.3f trees %.3f qgamma %.3f | " % (rep, tot, st["pretrace_s"], st["trees_s"], st["q_gamma_s"]) +
          " ".join("%s=%.3f" % (k, v) for k, v in acc.items()))
    del r
    torch.cuda.synchronize()
