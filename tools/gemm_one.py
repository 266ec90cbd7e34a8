"""Single tcgen05 NT GEMM shape, a few launches (ncu --set full target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_probe import run
M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (65536, 192, 192)))
run(M, N, K, tn=(len(sys.argv) > 4 and sys.argv[4] == "tn"), iters=4)
