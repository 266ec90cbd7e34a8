mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_dp.py tests/test_gpu_tokens_fused.py tests/test_gpu_live_reference.py -q -s > gpurun_out/r2/t06_dp.log 2>&1; tail -8 gpurun_out/r2/t06_dp.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/bench06_n2.log 2>&1; tail -c 600 gpurun_out/r2/bench06_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload qavitv2_c100 --batch 256 > gpurun_out/r2/bench06_qavitv2_b256_n2.log 2>&1; tail -c 600 gpurun_out/r2/bench06_qavitv2_b256_n2.log
python bench.py --gpus 1 --steps 20 --warmup 3 --workload qavitv2_c100 --batch 256 > gpurun_out/r2/bench06_qavitv2_b256_n1.log 2>&1; tail -c 400 gpurun_out/r2/bench06_qavitv2_b256_n1.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > gpurun_out/r2/bench06_n2_nograph.log 2>&1; tail -c 400 gpurun_out/r2/bench06_n2_nograph.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/profile_step.py --batch 8 --warmup 1 --dropout 0.1 --drop-path 0.1 > gpurun_out/r2/racecheck06.log 2>&1; tail -5 gpurun_out/r2/racecheck06.log
