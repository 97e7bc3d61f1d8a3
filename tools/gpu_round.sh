set -x
python -m pytest tests/test_gpu_widen.py -m gpu -x -q > gpurun_out/t_gpu_widen.log 2>&1
python tools/train_points.py > gpurun_out/train_points.log 2>&1
