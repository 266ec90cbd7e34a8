mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/t21_all.log 2>&1; tail -8 gpurun_out/r2/t21_all.log
