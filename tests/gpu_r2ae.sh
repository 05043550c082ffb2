#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2ae_pytest.log
grep -E "passed|failed|rror|assert" gpurun_out/r2ae_pytest.log | tail -5
(python tests/quick_ab_options.py --lanes 4 --reps 5 base: lt1:light_trace_mode=1
python tests/quick_ab_options.py --lanes 4 --reps 5 --fast base: lt1:light_trace_mode=1) 2>&1 | grep cfg | tee gpurun_out/r2ae_ab.log
python bench.py --steps 3 --warmup 3 --cpu-sample 65536 --no-equal-time --render-frames 48 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); sp=d['spcbpt']; print(sp['samples_per_s']/1e6, sp['fast_flavour']['samples_per_s']/1e6, sp['e2e'])"
