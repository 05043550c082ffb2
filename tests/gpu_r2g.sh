#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -s -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
grep -E "passed|failed|rror|assert" gpurun_out/r2g_pytest.log | tail -5
host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 8 --lanes 1 --no-images --quiet --save-state /tmp/st_ > /dev/null 2>&1
run() { # label, binary, lanes, extra...
  local label=$1 bin=$2 lanes=$3; shift 3
  for rep in 1 2 3; do
    host/_build/$bin --cache data/_ref/house.spcscene --dim=1920x1080 --frames 96 --lanes $lanes --no-images --quiet --no-pipeline --load-state /tmp/st_ "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$label lanes=$lanes rep=$rep ms_per_frame=%.3f launches=%d mean=%.5f' % (d['ms_per_frame'], d['kernel_launches'], d['image_mean']))"
  done
}
for lanes in 1 4; do
run exact spcbpt_render $lanes
run exact_sort spcbpt_render $lanes --option sort_hits=1
run exact_sort_tail32k spcbpt_render $lanes --option sort_hits=1 --option tail_threshold=32768
run fast_lt1 spcbpt_render_fast $lanes --option light_trace_mode=1
run fast_lt1_sort spcbpt_render_fast $lanes --option light_trace_mode=1 --option sort_hits=1
done 2>&1 | tee gpurun_out/r2g_ab.log
