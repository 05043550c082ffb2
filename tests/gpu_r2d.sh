#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
grep -E "passed|failed|rror|launches for|serial cores|vertices:" gpurun_out/r2d_pytest.log | tail -12
timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2d_bench.err
for lanes in 1 4; do for fl in "" _fast; do
timeout 300 host/_build/spcbpt_render$fl --cache data/_ref/house.spcscene --dim=1920x1080 --frames 48 --lanes $lanes --no-images --quiet --no-pipeline 2>&1 | tail -1 | cut -c1-330
done; done
for opt in "tail_threshold=-1" "light_trace_mode=1"; do
timeout 300 host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 48 --lanes 1 --no-images --quiet --no-pipeline --option $opt 2>&1 | tail -1 | cut -c1-330
done
timeout 300 host/_build/spcbpt_render_fast --cache data/_ref/house.spcscene --dim=1920x1080 --frames 48 --lanes 4 --no-images --quiet --option light_trace_mode=1 2>&1 | tail -1 | cut -c1-330
