#!/bin/bash
# round-2 GPU pass: tests, bench, ncu captures (large scene + house frame) exported as CSV on the box, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"
timeout 600 python bench.py --workload large --no-render --steps 5 > gpurun_out/r2a_bench_large.json 2> gpurun_out/r2a_bench_large.err; echo "bench large exit $?"
cap() {  # name, kernel regex, count, command...
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$rx" -c $cnt -o /tmp/$name -f "$@" > gpurun_out/${name}_ncu.log 2>&1; echo "ncu $name exit $?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
}
cap r2a_large k_trace_persist 7 python bench.py --workload large --no-render --steps 1 --warmup 1 --cpu-sample 1024
ncu -i /tmp/r2a_large.ncu-rep --page source --csv --kernel-name regex:k_trace_persist --launch-skip 5 --launch-count 1 > gpurun_out/r2a_large_B_source.csv 2>/dev/null
cap r2a_micro k_trace_persist 7 python bench.py --no-render --steps 1 --warmup 1 --cpu-sample 1024
cap r2a_house "k_trace_persist|k_eye_connect|k_eye_sample|k_eye_shade" 10 host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 1 --no-pipeline --no-images
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2a_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-render --cpu-sample 1024 > /dev/null 2>&1; echo "launch list exit $?"
du -sh gpurun_out; ls -la gpurun_out
