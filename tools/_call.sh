mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -x > gpurun_out/r2/t20_all.log 2>&1; tail -4 gpurun_out/r2/t20_all.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2/bench20.log 2>&1; tail -c 2500 gpurun_out/r2/bench20.log
