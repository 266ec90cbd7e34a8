"""Per-kernel time, DRAM bytes and achieved DRAM rate from an ncu csv taken with
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
(cold-cache, serialised launches: shares and byte counts are what to read, not absolute step time).
    python tools/step_metrics_summary.py profiles/r2_step_metrics_b4736.csv [top]"""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 70
K = collections.OrderedDict()
for r in csv.DictReader(lines):
    d = K.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]; m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    d[m] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0])
tot = byt = 0.0
for d in K.values():
    n = re.sub(r"\(.*", "", d["name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", ""))[:64]
    a = agg[n]
    t = d["gpu__time_duration.sum"]; b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    a[0] += 1; a[1] += t; a[2] += b; a[3] += d.get("smsp__inst_executed.sum", 0)
    a[4] += d.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0); a[5] = d.get("launch__registers_per_thread", 0)
    tot += t; byt += b
print(f"{len(K)} launches, sum of kernel durations {tot:.0f} us, DRAM traffic {byt / 1e9:.1f} GB ({byt / 6547.2e9 * 1e3:.1f} ms at the measured 6547 GB/s)")
print(f"{'kernel':64s} {'n':>4} {'us':>8} {'%':>5} {'GB':>6} {'TB/s':>5} {'Minst':>7} {'occ%':>4} regs")
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{n:64s} {a[0]:4d} {a[1]:8.0f} {100 * a[1] / tot:5.1f} {a[2] / 1e9:6.2f} {a[2] / a[1] / 1e6:5.2f} {a[3] / 1e6:7.1f} {a[4] / a[0]:4.0f} {a[5]:.0f}")
