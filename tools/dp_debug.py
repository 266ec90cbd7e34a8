"""2-rank debug of the overlapped bucketed all-reduce vs the cross-rank mean (torchrun --nproc-per-node 2 tools/dp_debug.py)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
import qavit_b200 as Q
from util import build_model
family = sys.argv[1] if len(sys.argv) > 1 else "qavitv2_c100"
Bl = 3
g = torch.Generator().manual_seed(77)
xg = torch.randn(world * Bl, 3, 32, 32, generator=g); yg = torch.randint(0, 100, (world * Bl,), generator=g)
x, y = xg[rank * Bl:(rank + 1) * Bl].to(dev), yg[rank * Bl:(rank + 1) * Bl].to(dev)
def fresh(overlap):
    model, ocfg, sd, _ = build_model(family, device=dev, precision="fp32")
    model.train()
    bank = [model.global_bank.global_k, model.global_bank.global_v]
    opt = Q.FusedAdamW(model.named_parameters(), lr=1e-3, betas=(0.9, 0.999), weight_decay=0.05, max_grad_norm=0.5, tail_elems=sum(p.numel() for p in bank))
    red = Q.GradAllReducer(opt, n_buckets=4, bank_params=bank, overlap=overlap)
    return model, opt, red
def run(overlap):
    model, opt, red = fresh(overlap)
    opt.zero_grad()
    if overlap: red.reset()
    loss = Q.cross_entropy(model(x), y, label_smoothing=0.1)
    loss.backward()
    return model, opt, red
m1, o1, r1 = run(False)
local = o1.flat_g.clone()
gath = [torch.empty_like(local) for _ in range(world)]
dist.all_gather(gath, local)
mean = torch.stack(gath).mean(0)
# second local run: how reproducible are the LOCAL gradients?
m1b, o1b, r1b = run(False)
torch.cuda.synchronize()
rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-30)).item()
if rank == 0: print("local grads run-to-run relerr", rel(o1b.flat_g, local))
m2, o2, r2 = run(True)
if rank == 0: print("pending after backward", r2._pending, "handles", len(r2._handles))
r2.finish()
torch.cuda.synchronize()
got = o2.flat_g * o2.grad_prescale
if rank == 0:
    print("overall relerr", rel(got, mean))
    offs = o2.seg_off.tolist()
    for bi, (lo, hi, buf) in enumerate(r2.buckets):
        print(f"bucket {bi}: params [{lo},{hi}) elems {offs[hi]-offs[lo]} relerr {rel(got[offs[lo]:offs[hi]], mean[offs[lo]:offs[hi]]):.3e} vs-local {rel(got[offs[lo]:offs[hi]], local[offs[lo]:offs[hi]]*1.0):.3e}")
    worst = sorted(((rel(got[offs[i]:offs[i+1]], mean[offs[i]:offs[i+1]]), n) for i, n in enumerate(o2.names) if mean[offs[i]:offs[i+1]].norm() > 0), reverse=True)[:12]
    for e, n in worst: print(f"  {e:.3e} {n}")
dist.destroy_process_group()
