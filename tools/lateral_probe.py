"""Times the HQAViT lateral CNN path (torch modules; scope row f-1) fwd+bwd under autocast: NCHW vs channels_last,
cudnn.benchmark on/off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qavit_b200 as Q

def run(cl, bench, B=1024):
    torch.backends.cudnn.benchmark = bench
    torch.manual_seed(0)
    m = Q.HQAViT(Q.HQAViTConfig(dropout=0.0, drop_path=0.0)).cuda().train()
    mods = [m.cnn_stem, m.lmfa2, m.lmfa3, m.lmfa4, m.rrcv2, m.rrcv3, m.rrcv4]
    if cl:
        for mm in mods:
            mm.to(memory_format=torch.channels_last)
    x = torch.randn(B, 3, 32, 32, device="cuda")
    if cl:
        x = x.contiguous(memory_format=torch.channels_last)
    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            f2, f3, f4 = m.cnn_stem(x)
            R = [m.rrcv2(m.lmfa2(f2), 8, 8), m.rrcv3(m.lmfa3(f3), 8, 8), m.rrcv4(m.lmfa4(f4), 8, 8)]
            loss = sum(r.float().square().mean() for r in R)
        loss.backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record(); torch.cuda.synchronize()
    print(f"channels_last={cl} cudnn.benchmark={bench}: {e0.elapsed_time(e1)/5:.2f} ms / step (cnn_stem+lmfa+rrcv fwd+bwd, B={B})")

for cl in (False, True):
    for bench in (False, True):
        run(cl, bench)
