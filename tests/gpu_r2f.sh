#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_pytest.log
grep -E "passed|failed|rror|launches for|serial cores|vertices:|assert" gpurun_out/r2f_pytest.log | tail -12
for t in "" "--host-trees"; do
host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 8 --lanes 1 --no-images --quiet $t 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('trees_s=%.4f pretrace_s=%.3f q_gamma_s=%.3f mean=%.6f' % (d['trees_s'], d['pretrace_s'], d['q_gamma_s'], d['image_mean']))"
done
