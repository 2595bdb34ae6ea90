# one N of the scaling set (round 2): bash tools/gpu_scaling.sh N  under  gpurun --gpus N  -- headline bench, TetraSpine population bench, multi-rank parity check
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_r02e_n$N.json 2> gpurun_out/bench_n$N.err; tail -c 600 gpurun_out/bench_r02e_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload tetraspine > gpurun_out/bench_r02e_tetraspine_n$N.json 2>> gpurun_out/bench_n$N.err; python -c "import json; d=json.load(open('gpurun_out/bench_r02e_tetraspine_n$N.json')); print('tetraspine n$N', d['value'], d['per_rank_kernel_ms'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/population_multirank_check.py 2>&1 | tail -3 | tee gpurun_out/population_multirank_n$N.txt
