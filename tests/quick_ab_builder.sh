python -m pytest tests/test_trace_gpu.py -x -q 2>&1 | tail -3
for b in lbvh ploc; do
SPC_BVH_BUILDER=$b python bench.py --no-render 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$b', 'value',round(d['value']), 'B ms',round(r['kernel_ms'],3), 'frac',round(r['frac'],3), 'nodes/tris',d['config']['per_set_nodes_tris_per_ray'], 'bvh_nodes', d['config']['bvh_nodes'])"
done
python - <<'PY'
import spcbpt_loader, time
pkg = spcbpt_loader.load()
import os
for b in ("ploc",):
    sc = pkg.scenes.heightfield_scene(708)
    ctx = pkg.Context(0); ctx.upload_scene(sc); print(b, ctx.bvh_stats())
    sc = pkg.scenes.load_spcscene("data/_ref/house.spcscene")
    ctx = pkg.Context(0); ctx.upload_scene(sc); print(b, 'house', ctx.bvh_stats())
PY
SPC_BVH_BUILDER=lbvh python - <<'PY'
import spcbpt_loader, time
pkg = spcbpt_loader.load()
sc = pkg.scenes.heightfield_scene(708)
ctx = pkg.Context(0); ctx.upload_scene(sc); print('lbvh', ctx.bvh_stats())
sc = pkg.scenes.load_spcscene("data/_ref/house.spcscene")
ctx = pkg.Context(0); ctx.upload_scene(sc); print('lbvh house', ctx.bvh_stats())
PY
