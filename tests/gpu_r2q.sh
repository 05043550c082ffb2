#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
python bench.py > gpurun_out/r2q_bench_1gpu.json 2> gpurun_out/r2q_bench_1gpu.err
echo "bench exit $? in $(( $(date +%s) - t0 )) s"
t0=$(date +%s)
python bench.py --impl reference > gpurun_out/r2q_bench_reference.json 2> gpurun_out/r2q_bench_reference.err
echo "reference arm exit $? in $(( $(date +%s) - t0 )) s"
tail -c 600 gpurun_out/r2q_bench_reference.json
python -c "
import json
d=json.load(open('gpurun_out/r2q_bench_1gpu.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['spcbpt']['samples_per_s'], d['spcbpt']['fast_flavour']['samples_per_s'], d['spcbpt'].get('reference_gpu',{}).get('samples_per_s'))
"
