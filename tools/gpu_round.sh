set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_r1.log 2>&1
python tools/determinism_check.py 500 > gpurun_out/determinism.log 2>&1
UTTT_SLOT_MODE=0 python tools/determinism_check.py 500 >> gpurun_out/determinism.log 2>&1
python tools/determinism_check.py 440 >> gpurun_out/determinism.log 2>&1
