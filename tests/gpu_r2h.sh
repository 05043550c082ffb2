#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_tree_build_gpu.py -m gpu -q -s -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2h_pytest.log
grep -E "passed|failed|rror|assert" gpurun_out/r2h_pytest.log | tail -5
for t in "" "--host-trees"; do
host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 8 --lanes 1 --no-images --quiet $t 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('trees_s=%.4f pretrace_s=%.3f q_gamma_s=%.3f mean=%.6f' % (d['trees_s'], d['pretrace_s'], d['q_gamma_s'], d['image_mean']))"
done
(for lanes in 1 4; do
python tests/quick_ab_options.py --lanes $lanes base: sort:sort_hits=1 notail:tail_threshold=-1 tail32k:tail_threshold=32768 tail512k:tail_threshold=524288 lt1:light_trace_mode=1 lt1sort:light_trace_mode=1,sort_hits=1
python tests/quick_ab_options.py --lanes $lanes --fast base: sort:sort_hits=1 lt1:light_trace_mode=1 lt1sort:light_trace_mode=1,sort_hits=1
done) 2>&1 | grep cfg | tee gpurun_out/r2h_ab.log
