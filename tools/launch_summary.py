"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: share of the step per kernel."""
import collections, csv, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if v != v:            # ncu occasionally reports one launch as nan
        continue
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = row["Kernel Name"].replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"at::native::", "at::", name)
    name = name[:110]
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
steps = max(1, next((a[0] for k, a in agg.items() if k.endswith("::ce_kernel")), 1))   # one cross-entropy launch per step
print(f"{steps} step(s): {tot / steps:.1f} us of kernel time and {sum(a[0] for a in agg.values()) / steps:.0f} launches per step (serialised, cold cache)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{t / steps:10.1f} us {100 * t / tot:5.1f}%  n={n / steps:6.1f} avg={t / n:8.1f}  {k}")
