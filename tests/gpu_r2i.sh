#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2i_bench_1gpu.json 2> gpurun_out/r2i_bench_1gpu.err; echo "bench exit $?"; tail -2 gpurun_out/r2i_bench_1gpu.err
timeout 900 python bench.py --workload large --no-equal-time > gpurun_out/r2i_bench_large.json 2> gpurun_out/r2i_bench_large.err; echo "bench large exit $?"; tail -2 gpurun_out/r2i_bench_large.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2i_launches_house.csv host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 3 --no-pipeline --no-images --quiet --load-state /tmp/none 2>/dev/null || true
host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 2 --lanes 1 --no-images --quiet --save-state /tmp/st_ > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2i_launches_house.csv host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 3 --no-pipeline --no-images --quiet --load-state /tmp/st_ > /dev/null 2>&1; echo "launch list exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2i_launches_house_fast_lt1.csv host/_build/spcbpt_render_fast --cache data/_ref/house.spcscene --dim=1920x1080 --frames 3 --no-pipeline --no-images --quiet --load-state /tmp/st_ --option light_trace_mode=1 > /dev/null 2>&1; echo "launch list fast exit $?"
timeout 600 ncu --set full --clock-control none -k regex:"k_eye_tail|k_light_trace_paths|k_eye_connect|k_eye_shade|k_eye_sample" -c 12 -o /tmp/r2i_house -f host/_build/spcbpt_render_fast --cache data/_ref/house.spcscene --dim=1920x1080 --frames 1 --no-pipeline --no-images --quiet --load-state /tmp/st_ --option light_trace_mode=1 > gpurun_out/r2i_ncu_house.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/r2i_house.ncu-rep --page raw --csv > gpurun_out/r2i_house_fast_raw.csv 2>/dev/null
du -sh gpurun_out
