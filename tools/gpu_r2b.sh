#!/bin/bash
# round 2, second GPU pass: new API / gating / replay tests, then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_widen.py -x -q -s > gpurun_out/r2b_api.log 2>&1; echo "api rc=$?" >> gpurun_out/r2b_api.log
timeout 1500 python -m pytest tests/test_gpu_replay.py -q -s > gpurun_out/r2b_replay.log 2>&1; echo "replay rc=$?" >> gpurun_out/r2b_replay.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2b_gpu_all.log 2>&1; echo "all rc=$?" >> gpurun_out/r2b_gpu_all.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2b_smoke.log
tail -n 3 gpurun_out/r2b_api.log gpurun_out/r2b_replay.log gpurun_out/r2b_gpu_all.log gpurun_out/r2b_smoke.log
