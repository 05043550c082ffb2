# A/B of the two binary-topology builders on the SPCBPT section of bench.py (house scene, 1080p)
for b in lbvh ploc; do
SPC_BVH_BUILDER=$b python bench.py --steps 5 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); s=d['spcbpt']; print('$b', 'Msamples/s',round(s['samples_per_s']/1e6,1), 'ms/frame',round(s['ms_per_frame'],3), 'pre_s', round(s['preprocess_s'],2), 'mean', s['image_mean'])"
done
