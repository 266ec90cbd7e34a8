# scratch command file for `gpurun -- 'bash tools/_call.sh'` (edited per experiment)
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
ms() { tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1; }
echo "lateral path on the main stream (QAVIT_LATERAL_SERIAL=1):"; QAVIT_LATERAL_SERIAL=1 $B 2>&1 | ms
echo "lateral path on its side stream (default):"; $B 2>&1 | ms
