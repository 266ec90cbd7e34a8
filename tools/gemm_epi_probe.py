"""tcgen05 NT GEMM with the fused epilogues (GELU dual output, bf16 residual, GELU-backward multiply) at the lateral
path's shapes: CUDA-event timing with L2 flushed; also the target of ncu --set full captures (argv: M N K mode)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qavit_b200 import _lib as L


def run(M, N, K, mode, iters=6):
    A = torch.randn(M, K, device="cuda").bfloat16()
    Wb = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.zeros(N, device="cuda")
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    C2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    aux = torch.randn(M, N, device="cuda").bfloat16()
    flush = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.qavit_test_gemm_epi(A.data_ptr(), K, M, N, K, Wb.data_ptr(), bias.data_ptr(), C.data_ptr(), C2.data_ptr(), mode,
                                          aux.data_ptr(), s))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    byt = (M * K + N * K + M * N * (1, 2, 2, 2)[mode]) * 2
    print(f"NT mode {mode} M={M} N={N} K={K}: {t:8.1f} us  {2.0 * M * N * K / t / 1e6:8.1f} TFLOP/s  {byt / t / 1e3:8.1f} GB/s (alg)")


if __name__ == "__main__":
    if len(sys.argv) >= 5:
        run(*(int(v) for v in sys.argv[1:5]), iters=3)
    else:
        for mode in (0, 1, 3):
            run(262144, 1024, 256, mode)
        run(262144, 256, 1024, 2)
        run(262144, 256, 1024, 0)
        for mode in (0, 1, 2, 3):
            run(65536, 192, 192, mode)
        run(65536, 96, 192, 1)
