"""Markdown table of the bench lines of a round: python tools/bench_summary.py gpurun_out/r2/bench31_*.log > profiles/r2_bench_summary.md"""
import json, os, sys
rows = []
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
    except Exception:
        continue
    if "unavailable" in d:
        continue
    cfg = d.get("config", {})
    rows.append((os.path.basename(p), d.get("impl", "ours"), cfg.get("workload", "?"), cfg.get("per_gpu_batch"), d.get("n_gpus"), d.get("value"), d.get("ms_per_step"),
                 (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("step_frac_of_sustained_peak"), d.get("gpu_launches")))
print("| file | impl | workload | per-GPU batch | GPUs | images/s | ms/step | e2e images/s | F_min share of sustained bf16 peak | launches/step |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    f = lambda v, fmt: "" if v is None else format(v, fmt)
    print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]} | {f(r[5], ',.0f')} | {f(r[6], '.2f')} | {f(r[7], ',.0f')} | {f(r[8], '.3f')} | {r[9] if r[9] is not None else ''} |")
