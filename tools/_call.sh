mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -x > gpurun_out/r2/t10_all.log 2>&1; tail -4 gpurun_out/r2/t10_all.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2/bench10.log 2>&1; tail -c 1800 gpurun_out/r2/bench10.log | grep -o '"ms_per_step": [0-9.]*' | head -1
