#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
grep -E "passed|failed|error|Gamma:|fast vs exact|reference-on-GPU|SPCBPT 128" gpurun_out/r2c_pytest.log | tail -12
timeout 600 host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 48 --lanes 4 --no-images --quiet > gpurun_out/r2c_cpp_1rank.json 2> gpurun_out/r2c_cpp_1rank.err; echo "cpp 1 rank exit $?"; cat gpurun_out/r2c_cpp_1rank.json
timeout 600 host/_build/spcbpt_render --cache data/_ref/house.spcscene --dim=1920x1080 --frames 48 --lanes 4 --ranks 2 --no-images --quiet > gpurun_out/r2c_cpp_2rank.json 2> gpurun_out/r2c_cpp_2rank.err; echo "cpp 2 ranks exit $?"; cat gpurun_out/r2c_cpp_2rank.json; tail -3 gpurun_out/r2c_cpp_2rank.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err; echo "bench 2 gpu exit $?"; tail -3 gpurun_out/r2c_bench_2gpu.err
