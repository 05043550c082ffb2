#!/bin/bash
# ad-hoc GPU sweep of the traversal launch knobs (not a test)
run() { timeout 300 python bench.py --no-render --steps 8 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$1', 'value %.0f  B %.3f ms  %.0f Mrays/s  frac %.3f' % (d['value'], r['kernel_ms'], r['mrays_per_s'], r['frac']))"; }
for t in 2 4 8 12 16; do SPC_FETCH_THRESHOLD=$t run "fetch=$t"; done
for p in 2 3 4 8 1000; do SPC_POSTPONE_DIV=$p run "postpone_div=$p"; done
for b in 6 7 8 10; do SPC_TRACE_BLOCKS_PER_SM=$b run "blocks=$b"; done
