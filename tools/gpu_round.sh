set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_r1.log 2>&1
UTTT_PROFILE=2 python tools/prof_selfplay.py --games 500 --reps 4 > gpurun_out/plain_p2.log 2>&1
UTTT_PROFILE=1 python tools/prof_selfplay.py --games 500 --reps 4 > gpurun_out/plain_p1.log 2>&1
python tools/config_points.py > gpurun_out/config_points.log 2>&1
