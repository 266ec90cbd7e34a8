for i in 1 2; do python -m pytest tests/test_gpu_live_reference.py -m gpu -q -s -k "hqavit_c100 or hqavit_stl96" 2>&1 | grep "three bf16\|passed\|failed"; done
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
