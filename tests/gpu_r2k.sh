#!/bin/bash
mkdir -p gpurun_out
(for lanes in 2 3 4 5 6 8; do
python tests/quick_ab_options.py --lanes $lanes --fast --reps 5 lt1:light_trace_mode=1
python tests/quick_ab_options.py --lanes $lanes --reps 5 base:
done) 2>&1 | grep cfg | tee gpurun_out/r2k_lanes.log
