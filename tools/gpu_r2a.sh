#!/bin/bash
# round 2, first GPU pass: bf16x3 numerics + replay tests, then the whole GPU suite and short bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_net.py -x -q -s -k "x3 or random_init" > gpurun_out/r2a_x3.log 2>&1; echo "x3 rc=$?" >> gpurun_out/r2a_x3.log
timeout 1500 python -m pytest tests/test_gpu_replay.py -x -q -s > gpurun_out/r2a_replay.log 2>&1; echo "replay rc=$?" >> gpurun_out/r2a_replay.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2a_gpu_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2a_gpu_all.log
timeout 600 python bench.py --steps 3 --warmup 3 --numerics bf16x3 --no-cpu-baseline --saturated-games 0 > gpurun_out/r2a_bench_x3.json 2> gpurun_out/r2a_bench_x3.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_x3.log gpurun_out/r2a_replay.log gpurun_out/r2a_gpu_all.log
