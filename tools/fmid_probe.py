"""CCF-FFN mid-section (ffn_mid.cu) timing in isolation: python tools/fmid_probe.py [B] [C] [side]  (QAVIT_LIB selects an A/B build)"""
import ctypes as Ct, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qavit_b200 import _lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
C = int(sys.argv[2]) if len(sys.argv) > 2 else 96
side = int(sys.argv[3]) if len(sys.argv) > 3 else 4
T = side * side
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(1 << 29, dtype=torch.uint8, device="cuda")
rn = lambda *sh: torch.randn(*sh, device="cuda")
h_pre, d_hn2 = rn(B, T, C).bfloat16(), rn(B, T, C).bfloat16()
par = [1 + 0.1 * rn(C), 0.1 * rn(C), rn(C, 3, 3) / 3, None, 1 + 0.1 * rn(C), 1 + 0.1 * rn(C), 0.1 * rn(C)]
hn2, d_hpre = torch.empty_like(h_pre), torch.empty_like(h_pre)
st1, st2 = torch.empty(B * T, 2, device="cuda"), torch.empty(B * T, 2, device="cuda")
gr = [torch.zeros(C, device="cuda") for _ in range(6)]
dw = torch.zeros(C, 9, device="cuda")
arr = lambda ts: (Ct.c_void_p * len(ts))(*[t.data_ptr() if t is not None else None for t in ts])
fi, fo = arr([h_pre] + par), arr([hn2, st1, st2])
bi, bo = arr([h_pre] + par + [d_hn2, st1, st2]), arr([d_hpre, gr[0], gr[1], dw, None, gr[2], gr[3], gr[4]])
for name, op, a, o, nbytes in (("fwd", 0, fi, fo, 2 * h_pre.numel() * 2), ("bwd", 1, bi, bo, 3 * h_pre.numel() * 2)):
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.qavit_test_ffn_mid(op, B, side, C, a, o, s))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    print(f"ffn_mid {name} B={B} C={C} side={side}: {t:7.1f} us   {nbytes / t / 1e3:7.1f} GB/s")
