set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_r1.log 2>&1
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
