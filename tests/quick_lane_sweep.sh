# frame-lane sweep on the shipped scene through the C++ driver: lanes x resident blocks of the persistent trace kernels
R="host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --no-images --quiet"
for rep in 1 2; do
for bps in ${BPS:-9 8 7}; do
for lanes in ${LANES:-2 3 4 6}; do
  SPC_TRACE_BLOCKS_PER_SM=$bps $R --frames ${FRAMES:-192} --lanes $lanes 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('bps $bps lanes', d['lanes'], 'ms/frame %.3f'%d['ms_per_frame'], 'Msamples/s %.1f'%(d['samples_per_s']/1e6))"
done
done
done
