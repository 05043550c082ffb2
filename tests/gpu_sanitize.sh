#!/bin/bash
# compute-sanitizer memcheck + racecheck on one small house render (training + frames, all options on) and on the traversal batches.
# Logs go to gpurun_out/ and are committed under profiles/ (VERDICT r1 item 9).
mkdir -p gpurun_out
ARGS="--cache data/_ref/house.spcscene --dim=256x144 --frames 2 --lanes 1 --no-images --quiet --no-pipeline --train-samples 20000 --q-samples 10000 --tree-samples 10000 --lt-cores 64 --lt-padding 400 --lt-per-core 50 --pretrace-cores 10000 --batch 10000"
for tool in memcheck racecheck; do
  for variant in "default" "--option light_trace_mode=1 --option sort_hits=1 --option tail_threshold=4096"; do
    tag=$(echo "$variant" | tr -c 'a-z0-9' '_' | cut -c1-24)
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 host/_build/spcbpt_render $ARGS $( [ "$variant" = default ] || echo $variant ) > gpurun_out/sanitize_${tool}_${tag}.log 2>&1
    echo "$tool [$variant] exit $?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${tag}.log | tail -1)"
  done
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_trace_gpu.py -m gpu -q -x > gpurun_out/sanitize_memcheck_trace_tests.log 2>&1
echo "memcheck trace tests exit $?: $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_trace_tests.log | tail -1) $(grep -E 'passed|failed' gpurun_out/sanitize_memcheck_trace_tests.log | tail -1)"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x -k "tile_partition" > gpurun_out/sanitize_memcheck_tile_partition.log 2>&1
echo "memcheck tile partition exit $?: $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_tile_partition.log | tail -1) $(grep -E 'passed|failed' gpurun_out/sanitize_memcheck_tile_partition.log | tail -1)"
