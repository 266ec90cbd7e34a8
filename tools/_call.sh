mkdir -p gpurun_out/r2
for v in v1 v2; do
QAVIT_LIB=$PWD/qa-vit_b200/libqavit_$v.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2/bench09_$v.log 2>&1; tail -c 1800 gpurun_out/r2/bench09_$v.log | grep -o '"ms_per_step": [0-9.]*' | head -1
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2/bench09_v0.log 2>&1; tail -c 1800 gpurun_out/r2/bench09_v0.log | grep -o '"ms_per_step": [0-9.]*' | head -1
QAVIT_PDL=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline > gpurun_out/r2/bench09_off.log 2>&1; tail -c 1800 gpurun_out/r2/bench09_off.log | grep -o '"ms_per_step": [0-9.]*' | head -1
