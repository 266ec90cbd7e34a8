mkdir -p gpurun_out/r2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 tools/dp_debug.py qavitv2_c100 2>&1 | grep -v "Warning\|warn\|\*\*\*\|OMP" | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29714 tools/dp_debug.py hqavit_c100 2>&1 | grep -v "Warning\|warn\|\*\*\*\|OMP" | tail -9
python -m pytest tests/test_gpu_dp.py -m gpu -q -s > gpurun_out/r2/t23_dp.log 2>&1; tail -3 gpurun_out/r2/t23_dp.log
