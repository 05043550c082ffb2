#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
grep -E "passed|failed|rror|assert" gpurun_out/r2n_pytest.log | tail -5
(for lanes in 1 4; do
python tests/quick_ab_options.py --lanes $lanes base: lt1:light_trace_mode=1
python tests/quick_ab_options.py --lanes $lanes --fast base: lt1:light_trace_mode=1
done) 2>&1 | grep cfg | tee gpurun_out/r2n_ab.log
