set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_r1.log 2>&1
for lv in 2 1; do UTTT_PROFILE=$lv python tools/prof_selfplay.py --games 500 --reps 4 > gpurun_out/plain_p$lv.log 2>&1; done
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
UTTT_TRUNK=3 ncu --set full --cache-control none --clock-control none --import-source on -k regex:heads_fc -s 3 -c 2 -f -o gpurun_out/prof_heads_r1 python tools/fwd_loop.py 500 6 > gpurun_out/ncu_heads.log 2>&1
