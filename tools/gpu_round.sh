# round-end evidence run (one GPU): tests, bench (both arms), forward-latency curve, ncu launch list + full captures
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_r1.log 2>&1
python bench.py --impl reference > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
for n in 16 64 148 222 296 370 400 444 500 518 592 740; do python tools/fwd_loop.py $n 200; done 2>&1 | grep "^n=" > gpurun_out/fwd_curve.log
UTTT_DEBUG_PHASES=1 python tools/pp_timeline.py 500 > gpurun_out/tl_pp.log 2>&1
UTTT_DEBUG_PHASES=1 python tools/trunk_timeline.py 345 > gpurun_out/tl.log 2>&1
UTTT_DEBUG_PHASES=1 python tools/trunk_timeline.py 40 > gpurun_out/tl40.log 2>&1
UTTT_PROFILE=2 UTTT_DEBUG_TREE=1 python tools/prof_selfplay.py --games 500 --reps 3 > gpurun_out/plain_r1.log 2>&1
python tools/determinism_check.py 500 > gpurun_out/determinism.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/launches_r1.csv python tools/prof_selfplay.py --games 500 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trunk_auto -s 200 -c 1 -f -o gpurun_out/prof_trunkpp_r1 python tools/prof_selfplay.py --games 500 > gpurun_out/ncu_pp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trunk_auto -s 3 -c 1 -f -o gpurun_out/prof_trunk2_r1 python tools/trunk_timeline.py 345 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tree_round -s 200 -c 1 -f -o gpurun_out/prof_tree_r1 python tools/prof_selfplay.py --games 500 > gpurun_out/ncu4.log 2>&1
UTTT_TRUNK=3 ncu --set full --cache-control none --clock-control none --import-source on -k regex:heads_fc -s 3 -c 1 -f -o gpurun_out/prof_heads_r1 python tools/fwd_loop.py 500 6 > gpurun_out/ncu_heads.log 2>&1
