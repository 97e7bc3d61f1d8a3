#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_net.py -q -x -k "variant_boundaries or reproducible" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_replay.py -q -x 2>&1 | tail -5
for i in 1 2; do timeout 120 python tools/cycle_time.py 1 6; done
} > gpurun_out/r2f.log 2>&1
cat gpurun_out/r2f.log
