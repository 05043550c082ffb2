# node steps per triangle phase: microbench value and house-scene frame time (sequential driver) for three builds
run() { python bench.py --no-render --steps 10 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'value',round(d['value']), 'B ms',round(r['kernel_ms'],3), 'frac',round(r['frac'],3), d['config']['per_set_nodes_tris_per_ray']['C'])"; }
SPCBPT_LIB=$PWD/alt_lib/steps4.so run steps4
R="host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --no-images --quiet --frames 96"
for rep in 1 2 3; do
for v in base steps3 steps4; do
  mkdir -p /tmp/l_$v; cp alt_lib/$v.so /tmp/l_$v/libspcbpt_b200.so
  LD_LIBRARY_PATH=/tmp/l_$v $R 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v house ms/frame %.3f'%d['ms_per_frame'])"
done
done
