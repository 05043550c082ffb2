# A/B of library builds on the traversal microbench: usage  bash tests/quick_ab_bench.sh libA.so libB.so ...
for rep in 1 2; do
for lib in "$@"; do
SPCBPT_LIB=$PWD/$lib python bench.py --no-render --steps 10 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$lib', 'value',round(d['value']), 'B ms',round(r['kernel_ms'],3), 'frac',round(r['frac'],3))"
done
done
