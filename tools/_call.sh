python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropout.py tests/test_gpu_live_reference.py -m gpu -q -x 2>&1 | tail -1
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
ms() { tail -1 | grep -o '"ms_per_step": [0-9.]*\|"gpu_launches": [0-9]*' | head -3 | tr '\n' ' '; echo; }
echo "hqavit:"; $B --steps 10 2>&1 | ms
echo "qavitv2:"; $B --workload qavitv2_c100 2>&1 | ms
python tools/kineto_step.py --graph --top 14 --workload qavitv2_c100 2>&1 | grep "bwr\|wall"
python tools/kineto_step.py --graph --top 30 2>&1 | grep "bwr\|wall"
