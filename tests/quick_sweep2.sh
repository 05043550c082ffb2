# sweeps of the persistent traversal kernel on the microbench: node steps per triangle phase (two builds), fetch threshold, postpone ratio
run() { python bench.py --no-render --steps 10 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1', 'value',round(d['value']), 'B ms',round(r['kernel_ms'],3), 'frac',round(r['frac'],3))"; }
SPCBPT_LIB=$PWD/alt_lib/steps3.so run "steps3"
run "steps2 default(6,5)"
for f in 3 4 8 10 12; do SPC_FETCH_THRESHOLD=$f run "steps2 fetch=$f"; done
for p in 3 4 8 100; do SPC_POSTPONE_DIV=$p run "steps2 postpone=$p"; done
