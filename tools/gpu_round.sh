set -x
python -m pytest tests/test_gpu_net.py -m gpu -x -q > gpurun_out/t_gpu_net.log 2>&1
UTTT_DEBUG_PHASES=1 python tools/pp_timeline.py 500 2>&1 | grep -E "trunk_auto|heads FC" > gpurun_out/tl_pp.log
UTTT_PROFILE=1 python tools/prof_selfplay.py --games 500 --reps 4 > gpurun_out/plain_p1.log 2>&1
