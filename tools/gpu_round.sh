set -x
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/gputests_r02a.log 2>&1
python bench.py > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
python bench.py --impl reference --steps 3 > gpurun_out/bench_r02a_reference.json 2>> gpurun_out/bench_r02a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02a.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tgb_step_kernel -s 3 -c 1 -o gpurun_out/prof_r02a -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python bench.py --workload tetraspine --no-cpu-baseline > gpurun_out/bench_r02a_tetraspine.json 2>> gpurun_out/bench_r02a.err
python bench.py --workload crater --no-cpu-baseline > gpurun_out/bench_r02a_crater.json 2>> gpurun_out/bench_r02a.err
python bench.py --envs 65536 --no-cpu-baseline --steps 5 > gpurun_out/bench_r02a_65536.json 2>> gpurun_out/bench_r02a.err
tail -3 gpurun_out/gputests_r02a.log; cat gpurun_out/bench_r02a.json
