"""Fused TokenLearner / TokenUpMix kernels (tokens_fused.cu) in isolation at the headline shape, for CUDA-event timing and ncu captures.
    python tools/tok_probe.py [B] [reps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qavit_b200 import _lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
N, Cc = 64, 192
s = torch.cuda.current_stream().cuda_stream
rn = lambda *sh: torch.randn(*sh, device="cuda")
x, lw, lb, W, b = rn(B, N, Cc), 1 + 0.1 * rn(Cc), 0.1 * rn(Cc), 0.15 * rn(16, Cc), 0.1 * rn(16)
S, xc, Z, dxc, dx = torch.empty(B, N, 16, device="cuda"), torch.empty(B, 16, Cc, device="cuda"), torch.empty(B, N, 16, device="cuda"), rn(B, 16, Cc), torch.empty(B, N, Cc, device="cuda")
dW, db, dlw, dlb = torch.zeros_like(W), torch.zeros_like(b), torch.zeros_like(lw), torch.zeros_like(lb)
Wu, bu = 0.3 * rn(N, 16), 0.2 * rn(N)
out, stats, dout, dxc2, dWu = torch.empty(B, N, Cc, device="cuda"), torch.empty(B * N, 2, device="cuda"), rn(B, N, Cc), torch.empty(B, 16, Cc, device="cuda"), torch.zeros_like(Wu)
arr = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
ops = [("tlf_fwd", 0, [x, lw, lb, W, b], [S, xc, Z], 4 * (x.numel() + xc.numel() + 2 * S.numel())),
       ("tlf_bwd", 1, [x, S, dxc, lw, lb, W, Z], [dx, dW, db, dlw, dlb], 4 * (2 * x.numel() + xc.numel() + 2 * S.numel())),
       ("upf_fwd", 2, [xc, Wu, bu, lw, lb], [out, stats], 4 * (xc.numel() + out.numel())),
       ("upf_bwd", 3, [xc, dout, stats, Wu, bu, lw], [dxc2, dWu, dlw, dlb], 4 * (2 * xc.numel() + dout.numel()))]
flush = torch.empty(1 << 29, dtype=torch.uint8, device="cuda")
for name, op, ins, outs, nbytes in ops:
    a, o = arr(ins), arr(outs)
    ts = []
    for i in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.qavit_test_tokens_fused(op, B, N, Cc, a, o, s))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"{name} B={B}: {t:7.1f} us   {nbytes / t / 1e3:7.1f} GB/s (algorithmic)")
