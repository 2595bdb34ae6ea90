# end-of-session measurement set, one GPU (round 2): tests, bench + reference arm, launch list, ncu --set full of the step kernel, side workloads
set -x
R=r02e
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/gputests_$R.log 2>&1
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python bench.py --impl reference --steps 3 > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_$R.err
TGB_TAIL_SPLIT=0 python bench.py --no-cpu-baseline > gpurun_out/bench_${R}_nosplit.json 2>> gpurun_out/bench_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
TGB_TAIL_SPLIT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tgb_step_kernel -s 3 -c 1 -o gpurun_out/prof_$R -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tgb_step_kernel -s 3 -c 1 -o gpurun_out/prof_${R}_tetra -f python bench.py --workload tetraspine --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tetra.log 2>&1
python bench.py --workload tetraspine > gpurun_out/bench_${R}_tetraspine.json 2>> gpurun_out/bench_$R.err
python bench.py --workload crater --no-cpu-baseline > gpurun_out/bench_${R}_crater.json 2>> gpurun_out/bench_$R.err
python bench.py --workload prism --no-cpu-baseline > gpurun_out/bench_${R}_prism.json 2>> gpurun_out/bench_$R.err
python bench.py --envs 65536 --no-cpu-baseline --steps 5 > gpurun_out/bench_${R}_65536.json 2>> gpurun_out/bench_$R.err
python bench.py --envs 1024 --no-cpu-baseline > gpurun_out/bench_${R}_1024.json 2>> gpurun_out/bench_$R.err
tail -3 gpurun_out/gputests_$R.log; cat gpurun_out/bench_$R.json
