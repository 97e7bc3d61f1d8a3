#!/bin/bash
mkdir -p gpurun_out
{
for pdl in 0 1; do for lvl in 0 1; do UTTT_PDL=$pdl timeout 120 python tools/cycle_time.py $lvl 6; done; done
UTTT_PDL=1 timeout 300 python -m pytest tests/test_gpu_replay.py tests/test_gpu_mcts.py -q -x 2>&1 | tail -3
} > gpurun_out/r2e_pdl.log 2>&1
cat gpurun_out/r2e_pdl.log
