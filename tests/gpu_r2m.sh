#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --no-equal-time > gpurun_out/r2m_bench_8gpu.json 2> gpurun_out/r2m_bench_8gpu.err; echo "bench 8 exit $?"; tail -2 gpurun_out/r2m_bench_8gpu.err
timeout 600 $TR --master-port 29522 bench.py --gpus 8 --render-dim 3840x2160 --no-equal-time > gpurun_out/r2m_bench_8gpu_4k.json 2> gpurun_out/r2m_bench_8gpu_4k.err; echo "bench 8 4k exit $?"; tail -2 gpurun_out/r2m_bench_8gpu_4k.err
timeout 600 $TR --master-port 29523 bench.py --gpus 8 --workload large --no-equal-time > gpurun_out/r2m_bench_8gpu_large.json 2> gpurun_out/r2m_bench_8gpu_large.err; echo "bench 8 large exit $?"; tail -2 gpurun_out/r2m_bench_8gpu_large.err
timeout 300 host/_build/spcbpt_render_fast --cache data/_ref/house.spcscene --dim=3840x2160 --frames 32 --lanes 4 --ranks 8 --no-images --quiet --option light_trace_mode=1 > gpurun_out/r2m_cpp_8ranks_4k.json 2> gpurun_out/r2m_cpp_8ranks_4k.err; echo "cpp 8 ranks exit $?"; tail -1 gpurun_out/r2m_cpp_8ranks_4k.json | cut -c1-400
