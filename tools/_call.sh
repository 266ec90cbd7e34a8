python -m pytest tests/test_gpu_tokens_fused.py tests/test_gpu_parity.py tests/test_gpu_live_reference.py -m gpu -q -x 2>&1 | tail -1
python tools/tok_probe.py 4736 5
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
ms() { tail -1 | grep -o '"ms_per_step": [0-9.]*\|"gpu_launches": [0-9]*' | head -3 | tr '\n' ' '; echo; }
echo "step:"; $B 2>&1 | ms
echo "step:"; $B 2>&1 | ms
