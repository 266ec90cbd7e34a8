"""Per-phase timeline of one tcgen05 NT GEMM launch (diagnostic build: make EXTRA=-DQV_GEMM_TRACE BUILD=build_trace OUT=../libqavit_trace.so,
run with QAVIT_LIB=qa-vit_b200/libqavit_trace.so)."""
import ctypes, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qavit_b200 import _lib as L

NAMES = ["entry", "prefill issued", "set-up barrier passed", "MMA: first k-block landed", "MMA: tile 0 issued", "EPI: tile 0 accumulator ready",
         "EPI: tile 0 stores issued", "MMA: last tile issued", "EPI: last tile accumulator ready", "EPI: last tile stores issued",
         "EPI: stores drained", "exit", "first loads issued"]

def run(M, N, K, flush=True):
    A = torch.randn(M, K, device="cuda").bfloat16()
    Wb = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.zeros(N, device="cuda")
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = torch.empty(1 << 29, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    fn = L.lib.qavit_test_gemm_trace_read
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p]
    for it in range(3):
        if flush:
            fl.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.qavit_test_gemm_nt(1, A.data_ptr(), K, M, N, K, None, Wb.data_ptr(), bias.data_ptr(), C.data_ptr(), 0, s))
        e1.record()
        torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 128)()
    assert fn(ctypes.addressof(buf)) == 0
    print(f"NT M={M} N={N} K={K} flush={flush}: events {e0.elapsed_time(e1) * 1e3:.1f} us")
    g0 = buf[64]
    for which, base in (("first CTA", 0), ("last CTA", 32)):
        print(f"  {which}: (clock64 delta in cycles | globaltimer ns since first CTA's entry)")
        for i, n in enumerate(NAMES):
            print(f"    {n:36s} {buf[base + i] - buf[base]:8d} cyc   {buf[64 + base + i] - g0:8d} ns")

if __name__ == "__main__":
    for shape in [(75776, 576, 192), (75776, 192, 192)]:
        run(*shape)
        run(*shape, flush=False)
