"""One HQAViT CIFAR-100 bf16 training step inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qavit_b200 as Q

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--model", default="hqavit")
ap.add_argument("--dropout", type=float, default=0.0)
ap.add_argument("--drop-path", type=float, default=0.0)
a = ap.parse_args()
torch.manual_seed(42)
if a.model == "hqavit":
    model = Q.HQAViT(Q.HQAViTConfig(dropout=a.dropout, drop_path=a.drop_path))
    if a.dropout == 0.0:
        for n in ("fuse2", "fuse3", "fuse4"):
            getattr(model, n).cat_mlp[3].p = 0.0
elif a.model == "tinyin":
    model = Q.HQAViT(Q.HQAViTConfig(img_size=64, num_classes=200, depth=12, num_learned_tokens=64, dropout=a.dropout, drop_path=a.drop_path),
                     stage_depths=(2, 2, 6, 2), square_tokens=True)
else:
    model = Q.QAViT(Q.QAViTConfig(dropout=a.dropout, drop_path=a.drop_path))
model = model.cuda().train().set_precision("bf16")
opt = Q.FusedAdamW(model.named_parameters(), lr=6e-4, betas=(0.95, 0.999), weight_decay=0.06, max_grad_norm=0.5)
nograd = ("swa.norm.", "msda.norm.", "cga.norm.", "write_norm.", "write_compression.", "write_gate.")
opt.set_grad_mask([not any(s in n for s in nograd) for n, _ in model.named_parameters()])
S, NC = (64, 200) if a.model == "tinyin" else (32, 100)
x = torch.randn(a.batch, 3, S, S, device="cuda")
y = torch.randint(0, NC, (a.batch,), device="cuda")

def step():
    opt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(x)
    loss = Q.cross_entropy(logits, y, label_smoothing=0.12)
    loss.backward()
    opt.clip()
    opt.step()
    return loss

for _ in range(a.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
l = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", l.item())
