#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_replay.py tests/test_gpu_widen.py -q -x 2>&1 | tail -4
UTTT_PROFILE=2 timeout 120 python tools/prof_selfplay.py --games 500 --reps 3 2>&1 | grep "^games"
for i in 1 2; do timeout 120 python tools/cycle_time.py 1 6; done
} > gpurun_out/r2g.log 2>&1
cat gpurun_out/r2g.log
