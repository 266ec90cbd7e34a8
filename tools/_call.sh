QAVIT_LATERAL_SERIAL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1
python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-eager-baseline 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1
