python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
QAVIT_LIB=qa-vit_b200/libqavit_trace.so python tools/gemm_trace.py 2>&1 | grep -v "last CTA" | head -64 | awk 'NR<=16 || (NR>30 && NR<=46)'
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*' | head -1
