mkdir -p gpurun_out/r2
echo "cuda-core:"; QV_NO_DWT=1 python tools/dw_probe.py
echo "tensor-core:"; python tools/dw_probe.py
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
ms() { tail -1 | grep -o '"ms_per_step": [0-9.]*\|"gpu_launches": [0-9]*' | head -3 | tr '\n' ' '; echo; }
echo "step dwt:"; $B 2>&1 | ms
echo "step QV_NO_DWT=1:"; QV_NO_DWT=1 $B 2>&1 | ms
echo "step dwt:"; $B 2>&1 | ms
echo "step QV_NO_DWT=1:"; QV_NO_DWT=1 $B 2>&1 | ms
