# sweep of the SAH triangle cost of the wide-collapse DP on the traversal microbench and the house scene
for c in ${COSTS:-0.3 0.5 0.7 1.0 1.5}; do
SPC_BVH_COST_PRIM=$c python bench.py --no-render --steps 10 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('cost $c', 'value',round(d['value']), 'B ms',round(r['kernel_ms'],3), 'Mrays/s B', round(r['mrays_per_s']), 'nodes/tris',[[round(x,2) for x in v] for v in d['config']['per_set_nodes_tris_per_ray'].values()], 'bvh_nodes', d['config']['bvh_nodes'])"
SPC_BVH_COST_PRIM=$c host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --no-images --quiet --frames 96 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('   house lanes 1 ms/frame %.3f'%d['ms_per_frame'])"
done
