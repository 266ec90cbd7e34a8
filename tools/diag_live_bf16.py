"""Diagnostic (GPU box): where does the bf16 run of HQAViT leave the fp32 LIVE reference?  Switches one module family at a
time between fp32 and bf16 and prints logits / all-gradient distances, plus the gradient error per top-level module."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import qavit_b200 as Q
from baseline import live_reference as LR
from qavit_b200 import modules as M
from util import rel_l2, rel_max

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
import copy

B = int(os.environ.get("DIAG_B", "16"))
mod, ref = LR.build("HQAViT_CIFAR100")
state = {k: v.detach().clone() for k, v in ref.state_dict().items()}
g = torch.Generator().manual_seed(3)
x = torch.randn(B, 3, 32, 32, generator=g).cuda()
y = torch.randint(0, 100, (B,), generator=g).cuda()
crit = torch.nn.CrossEntropyLoss(label_smoothing=0.1)


def run_ref(autocast):
    m = copy.deepcopy(ref).cuda().train()
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        lo = m(x)
        loss = crit(lo, y)
    loss.backward()
    return lo.detach().float(), {n: p.grad.detach().float() for n, p in m.named_parameters() if p.grad is not None}


def run_ours(prec):
    """prec: dict family -> 'fp32' | 'bf16' for 'block', 'sf', 'patch', 'lateral'"""
    m = Q.HQAViT(ref.config)
    for n in ("fuse2", "fuse3", "fuse4"):
        getattr(m, n).cat_mlp[3].p = 0.0
    m.load_state_dict(state, strict=True)
    m = m.cuda().train()
    m.precision = prec["lateral"]
    for s in m.modules():
        if isinstance(s, M.QuadAttentionBlock):
            s.precision = prec["block"]
        elif isinstance(s, M.SplitFusion):
            s.precision = prec["sf"]
        elif isinstance(s, M.PatchEmbed):
            s.precision = prec["patch"]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lo = m(x)
    Q.cross_entropy(lo, y, label_smoothing=0.1).backward()
    return lo.detach().float(), {n: p.grad.detach().float() for n, p in m.named_parameters() if p.grad is not None}


def gerr(a, r, by_group=False):
    num = collections.defaultdict(float)
    den = collections.defaultdict(float)
    for n, gr in r.items():
        k = n.split(".")[0] if by_group else "all"
        num[k] += (a[n] - gr).norm().item() ** 2
        den[k] += gr.norm().item() ** 2
    if by_group:
        tot = sum(den.values())
        return {k: ((num[k] / den[k]) ** 0.5, (num[k] / tot) ** 0.5) for k in num}
    return (num["all"] / den["all"]) ** 0.5


r32, g32 = run_ref(False)
r16, g16 = run_ref(True)
print(f"reference autocast vs fp32: logits max {rel_max(r16, r32):.2e} l2 {rel_l2(r16, r32):.2e} grads {gerr(g16, g32):.2e}")
fams = ["block", "sf", "patch", "lateral"]
configs = [("all bf16", {f: "bf16" for f in fams}), ("all fp32", {f: "fp32" for f in fams})]
for f in fams:
    configs.append((f"only {f} bf16", {k: ("bf16" if k == f else "fp32") for k in fams}))
for f in fams:
    configs.append((f"only {f} fp32", {k: ("fp32" if k == f else "bf16") for k in fams}))
for name, prec in configs:
    lo, gr = run_ours(prec)
    print(f"{name:22s}: logits max {rel_max(lo, r32):.2e} l2 {rel_l2(lo, r32):.2e} grads {gerr(gr, g32):.2e}")
    if name in ("all bf16", "only block bf16"):
        grp = gerr(gr, g32, by_group=True)
        for k, (rel, share) in sorted(grp.items(), key=lambda kv: -kv[1][1])[:12]:
            print(f"      {k:18s} rel {rel:.2e}  share-of-total {share:.2e}")
