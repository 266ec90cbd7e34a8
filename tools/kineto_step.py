"""In-situ kernel durations of the headline training step (torch.profiler / CUPTI activity records: no replay, no cache flush, the
real stream concurrency) -- the complement of the ncu launch lists, whose per-kernel times are cold-cache and serialised.
    python tools/kineto_step.py [--graph] [--batch 4736] [--top 60]"""
import argparse, collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import qavit_b200 as Q
from torch.profiler import ProfilerActivity, profile

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4736)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--top", type=int, default=60)
ap.add_argument("--dropout", type=float, default=0.1)
ap.add_argument("--workload", default="hqavit_c100", help="one of bench.py's WORKLOADS")
a = ap.parse_args()
torch.manual_seed(42)
import bench
model, w = bench.build_workload(Q, a.workload, a.dropout, a.dropout)
model = model.cuda().train().set_precision("bf16")
opt = Q.FusedAdamW(model.named_parameters(), lr=6e-4, betas=(0.95, 0.999), weight_decay=0.06, max_grad_norm=0.5)
x = torch.randn(a.batch, 3, w["img"], w["img"], device="cuda")
y = torch.randint(0, w["classes"], (a.batch,), device="cuda")

def eager():
    opt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(x)
    loss = Q.cross_entropy(logits, y, label_smoothing=0.12)
    loss.backward()
    opt.clip()
    opt.step()

if a.graph:
    g = Q.GraphedTrainStep(model, opt, x, y, label_smoothing=0.12)
    step = lambda: g(x, y)
else:
    step = eager
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
t0, t1 = min(e.time_range.start for e in evs), max(e.time_range.end for e in evs)
for e in evs:
    n = re.sub(r"\(.*", "", e.name.replace("(anonymous namespace)::", "").replace("void ", ""))[:80]
    agg[n][0] += 1
    agg[n][1] += e.time_range.end - e.time_range.start
tot = sum(v[1] for v in agg.values())
print(f"2 steps: wall {t1 - t0:.0f} us, sum of kernel durations {tot:.0f} us over {sum(v[0] for v in agg.values())} records")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"{t / 2:10.1f} us {100 * t / tot:5.1f}%  n={n // 2:4d} avg={t / n:8.1f}  {k}")

# per-stream occupancy and the largest idle gaps of the busiest stream (what the critical path waits for)
by_stream = collections.defaultdict(list)
for e in evs:
    by_stream[e.device_resource_id if hasattr(e, "device_resource_id") else getattr(e, "stream", 0)].append(e)
print("streams:")
for sid, es in sorted(by_stream.items(), key=lambda kv: -sum(x.time_range.end - x.time_range.start for x in kv[1])):
    busy = sum(x.time_range.end - x.time_range.start for x in es)
    print(f"  stream {sid}: {len(es)} records, busy {busy / 2:.0f} us / step")
main = max(by_stream.values(), key=lambda es: sum(x.time_range.end - x.time_range.start for x in es))
main.sort(key=lambda e: e.time_range.start)
gaps = []
for a, b in zip(main, main[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 3:
        gaps.append((g, a.name[:60], b.name[:60]))
print(f"main stream: {len(gaps)} gaps > 3 us, total {sum(g[0] for g in gaps) / 2:.0f} us / step; all gaps {sum(max(0, b.time_range.start - a.time_range.end) for a, b in zip(main, main[1:])) / 2:.0f} us / step")
for g, a, b in sorted(gaps, reverse=True)[:25]:
    print(f"  {g:8.1f} us  after {a}  before {b}")
