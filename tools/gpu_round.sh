set -x
python -m pytest tests/test_gpu_mcts.py -m gpu -x -q > gpurun_out/t_gpu_mcts.log 2>&1
UTTT_DEBUG_TREE=1 UTTT_PROFILE=2 python tools/prof_selfplay.py --games 500 --reps 3 > gpurun_out/tree_dbg.log 2>&1
UTTT_PROFILE=1 python tools/prof_selfplay.py --games 500 --reps 4 > gpurun_out/plain_p1.log 2>&1
