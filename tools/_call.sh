mkdir -p gpurun_out/r2
for i in 1 2 3; do python -m pytest tests/test_gpu_live_reference.py -m gpu -q -s -k "hqavit_stl96 or hqavit_c100" 2>&1 | grep "^live\|passed\|failed" | sed -E 's/.*bf16 ours-vs-ref-fp32 (logits [0-9.e-]+) \(l2 [0-9.e-]+\) (grads [0-9.e-]+).*/\1 \2/' ; done
echo "--- QV_NO_XSTK=1"; QV_NO_XSTK=1 python -m pytest tests/test_gpu_live_reference.py -m gpu -q -s -k "hqavit_stl96" 2>&1 | grep "^live\|passed\|failed" | sed -E 's/.*bf16 ours-vs-ref-fp32 (logits [0-9.e-]+) \(l2 [0-9.e-]+\) (grads [0-9.e-]+).*/\1 \2/'
python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py tests/test_gpu_dropout.py -m gpu -q -x 2>&1 | tail -2
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline"
ms() { tail -1 | grep -o '"ms_per_step": [0-9.]*\|"gpu_launches": [0-9]*' | head -3 | tr '\n' ' '; echo; }
echo "stacked projections:"; $B 2>&1 | ms
echo "QV_NO_XSTK=1:"; QV_NO_XSTK=1 $B 2>&1 | ms
echo "qavitv2 stacked:"; $B --workload qavitv2_c100 2>&1 | ms
echo "qavitv2 QV_NO_XSTK=1:"; QV_NO_XSTK=1 $B --workload qavitv2_c100 2>&1 | ms
