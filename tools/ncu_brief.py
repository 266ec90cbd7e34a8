"""Compact per-kernel table from an .ncu-rep (ncu --page raw --csv): time, DRAM traffic, issue activity, top stalls."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
def g(d, name):
    return d[hdr.index(name)] if name in hdr else "nan"
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for d in data:
    name = g(d, "Kernel Name").replace("void ", "").replace("<unnamed>::", "")[:58]
    t = float(g(d, "gpu__time_duration.sum"))
    rd, wr = g(d, "dram__bytes_read.sum"), g(d, "dram__bytes_write.sum")
    st = sorted(((float(d[hdr.index(s)]), s[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for s in stalls), reverse=True)[:4]
    print(f"{name:58s} grid {g(d, 'Grid Size'):>13} {t:8.1f}us rd {rd[:7]:>7} wr {wr[:7]:>7} regs {g(d, 'launch__registers_per_thread'):>3} "
          f"occ {float(g(d, 'sm__warps_active.avg.pct_of_peak_sustained_active')):4.0f}% issue {float(g(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active')):4.0f}% "
          f"inst {float(g(d, 'smsp__inst_executed.sum')) / 1e6:7.1f}M dram {float(g(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):4.0f}% "
          + " ".join(f"{n}={v:.1f}" for v, n in st))
