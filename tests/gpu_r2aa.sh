#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2aa_pytest.log
grep -E "passed|failed|rror|assert" gpurun_out/r2aa_pytest.log | tail -5
(python tests/quick_ab_options.py --lanes 1 --reps 5 base: lt1:light_trace_mode=1
python tests/quick_ab_options.py --lanes 1 --reps 5 --fast base: lt1:light_trace_mode=1
python tests/quick_ab_options.py --lanes 4 --reps 5 --fast base: lt1:light_trace_mode=1) 2>&1 | grep cfg | tee gpurun_out/r2aa_ab.log
