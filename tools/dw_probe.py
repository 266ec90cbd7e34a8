"""Depthwise 8 x 8 forward timing (bf16): QV_NO_DWT=1 selects the CUDA-core kernel.  python tools/dw_probe.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qavit_b200 import _lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(1 << 29, dtype=torch.uint8, device="cuda")
for C in (64, 128, 256):
    for K in (3, 5, 7):
        x = torch.randn(B, 8, 8, C, device="cuda").bfloat16()
        w = torch.randn(C, K, K, device="cuda"); b = torch.randn(C, device="cuda")
        y = torch.empty_like(x)
        ts = []
        for i in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(L.lib.qavit_dwconv_forward(x.data_ptr(), 1, B, 8, 8, C, K, w.data_ptr(), b.data_ptr(), y.data_ptr(), s))
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[len(ts) // 2]
        print(f"C={C:3d} K={K}: {t:7.1f} us   {2 * x.numel() * 2 / t / 1e3:7.1f} GB/s")
