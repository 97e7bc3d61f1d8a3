#!/bin/bash
# round 2, third GPU pass: whole GPU suite, full bench line, rules points (plain), config-5 cycle on 1 GPU
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c_gpu_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2c_gpu_all.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?" >> gpurun_out/r2c_bench.err
timeout 300 python tools/prof_rules.py 22 > gpurun_out/r2c_rules_plain.log 2>&1
tail -n 5 gpurun_out/r2c_gpu_all.log gpurun_out/r2c_bench.err
