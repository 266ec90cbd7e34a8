mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -x -k "dp or nccl or rank" 2>&1 | tail -3 | tee gpurun_out/r2/t41_dp.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/bench41_n2.log 2>&1; tail -1 gpurun_out/r2/bench41_n2.log | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload qavitv2_c100 --batch 256 > gpurun_out/r2/bench41_qavitv2_b256_n2.log 2>&1; tail -1 gpurun_out/r2/bench41_qavitv2_b256_n2.log | cut -c1-400
