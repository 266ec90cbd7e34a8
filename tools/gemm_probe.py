"""Runs the tcgen05 NT / TN GEMMs at the block's dominant shapes (for ncu --set full captures and CUDA-event timing)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qavit_b200 import _lib as L

def run(M, N, K, tn=False, iters=6):
    A = torch.randn(M, K, device="cuda").bfloat16()
    W = torch.randn(N, K, device="cuda") / math.sqrt(K)
    Wb = W.bfloat16()
    bias = torch.zeros(N, device="cuda")
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    dW = torch.zeros(N, K, device="cuda")
    flush = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if tn:
            L.check(L.lib.qavit_test_gemm_tn(1, C.data_ptr(), N, A.data_ptr(), K, M, N, K, dW.data_ptr(), None, s))
        else:
            L.check(L.lib.qavit_test_gemm_nt(1, A.data_ptr(), K, M, N, K, W.data_ptr(), Wb.data_ptr(), bias.data_ptr(), C.data_ptr(), 0, s))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    byt = (M * K + N * K + M * N) * 2
    print(f"{'TN' if tn else 'NT'} M={M} N={N} K={K}: {t:8.1f} us  {2.0 * M * N * K / t / 1e6:8.1f} TFLOP/s  {byt / t / 1e3:8.1f} GB/s (alg)")

if __name__ == "__main__":
    for (M, N, K) in [(65536, 576, 192), (65536, 192, 192), (16384, 576, 192), (16384, 192, 192), (65536, 96, 192), (65536, 192, 96), (65536, 48, 192), (262144, 192, 192)]:
        run(M, N, K)
    for (M, N, K) in [(65536, 576, 192), (65536, 192, 192), (16384, 192, 192), (65536, 96, 192)]:
        run(M, N, K, tn=True)
