#!/bin/bash
# final-code ncu evidence of a house frame: launch lists (exact default / fast + lt1) and --set full of every per-frame kernel at bounce 0/1
mkdir -p gpurun_out
host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 2 --lanes 1 --no-images --quiet --save-state /tmp/st_ > /dev/null 2>&1
R="--cache data/_ref/house.spcscene --dim=1920x1080 --no-pipeline --no-images --quiet --load-state /tmp/st_"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2v_launches_house.csv host/_build/spcbpt_render $R --frames 3 > /dev/null 2>&1; echo "launch list exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2v_launches_house_fast_lt1.csv host/_build/spcbpt_render_fast $R --frames 3 --option light_trace_mode=1 > /dev/null 2>&1; echo "launch list fast exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"k_trace_persist|k_eye_tail|k_light_trace|k_eye_connect|k_eye_shade|k_eye_sample|k_lvc|k_lt_" -c 40 -o /tmp/r2v_house -f host/_build/spcbpt_render_fast $R --frames 1 --option light_trace_mode=1 > gpurun_out/r2v_ncu_house.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/r2v_house.ncu-rep --page raw --csv > gpurun_out/r2v_house_fast_raw.csv 2>/dev/null
du -sh gpurun_out
