# scratch command file for `gpurun -- 'bash tools/_call.sh'` (edited per experiment)
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/r2/t44_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --workload qavitv2_c100 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>&1 | tail -1 | cut -c1-230
