# A/B of two builds of the library through the C++ host driver on the shipped scene: alt_lib/ (baseline) vs the tree's build.
# usage: bash tests/quick_ab_lib.sh [lanes...]
R="host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --no-images --quiet"
for rep in 1 2; do
for lanes in ${@:-1 3 4}; do
  for which in alt new; do
    if [ $which = alt ]; then export LD_LIBRARY_PATH=$PWD/alt_lib; else unset LD_LIBRARY_PATH; fi
    $R --frames ${FRAMES:-192} --lanes $lanes 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$which lanes', d['lanes'], 'ms/frame %.3f'%d['ms_per_frame'], 'Msamples/s %.1f'%(d['samples_per_s']/1e6), 'mean', d['image_mean'])"
  done
done
done
