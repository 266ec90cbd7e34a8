"""One HQAViT CIFAR-100 bf16 training step with the CUDA profiler range open only around ONE instance of every distinct piece of the
step -- one wrapped quad block (forward + backward), one SplitFusion, the whole lateral path, patch embed, head, loss, clip + AdamW --
so that an `ncu --profile-from-start off` capture holds every kernel of the step a few times instead of ~950 launches.
    ncu --section ... --profile-from-start off -o rep python tools/profile_parts.py --batch 1184"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qavit_b200 as Q
import qavit_b200.functional as QF

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1184)
ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--variant", default="v1")
a = ap.parse_args()
torch.manual_seed(42)
model = Q.HQAViT(Q.HQAViTConfig(dropout=a.dropout, drop_path=a.dropout), variant=a.variant).cuda().train().set_precision("bf16")
opt = Q.FusedAdamW(model.named_parameters(), lr=6e-4, betas=(0.95, 0.999), weight_decay=0.06, max_grad_norm=0.5)
x = torch.randn(a.batch, 3, 32, 32, device="cuda")
y = torch.randint(0, 100, (a.batch,), device="cuda")
armed = [False]
def start(*_):
    if armed[0]:
        torch.cuda.profiler.start()      # (hooks must return None: a value would replace the module's input / output)


def stop(*_):
    if armed[0]:
        torch.cuda.profiler.stop()


for m in (model.stage2_blocks[1], model.fuse3, model.patch_embed):
    m.register_forward_pre_hook(start)
    m.register_forward_hook(stop)
    m.register_full_backward_pre_hook(start)
    m.register_full_backward_hook(stop)
for name in ("lateral", "lateral_adapter"):
    f = getattr(model, name)
    def wrap(f=f):
        def g(*args, **kw):
            start()
            try:
                return f(*args, **kw)
            finally:
                stop()
        return g
    setattr(model, name, wrap())
_bwd = QF.LateralPartFn.backward
def lat_bwd(ctx, dout):
    start()
    try:
        return _bwd(ctx, dout)
    finally:
        stop()
QF.LateralPartFn.backward = staticmethod(lat_bwd)

def step():
    opt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(x)                      # (the head is the last launch group of forward: opened below)
    start()
    loss = Q.cross_entropy(logits, y, label_smoothing=0.12)
    stop()
    loss.backward()
    start()
    opt.clip()
    opt.step()
    stop()
    return loss

for _ in range(2):
    step()
torch.cuda.synchronize()
armed[0] = True
l = step()
torch.cuda.synchronize()
print("loss", l.item())
