"""Opcode census of libqavit_b200.so per kernel (cuobjdump -sass): which kernels carry tcgen05 (UTC*MMA), TMA (UTMALDG / UTMASTG),
tensor-memory loads (LDTM), legacy tensor-core MMA (HMMA), ldmatrix (LDSM), cp.async (LDGSTS), atomics / reductions.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "qa-vit_b200", "libqavit_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = {}
names = sorted(set(re.findall(r"Function : (\S+)", out)))
if names:
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "LDSM", "LDGSTS", "ATOM", "RED", "BAR", "SHFL", "MUFU"]
cur, stats = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        stats[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        stats[cur]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                stats[cur][k] += 1
print(f"# {os.path.relpath(so, ROOT)}: {len(stats)} kernels; columns = SASS instruction counts (static), sm_100a")
print(f"{'kernel':88s} {'total':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
tot = collections.Counter()
for fn, c in sorted(stats.items(), key=lambda kv: demangle.get(kv[0], kv[0])):
    name = demangle.get(fn, fn)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    depth, cut = 0, len(name)
    for i, ch in enumerate(name):          # drop the parameter list, keep template arguments such as <(int)1>
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    name = name[:cut].replace("void ", "")[:88]
    print(f"{name:88s} {c['total']:7d} " + " ".join(f"{c[k]:7d}" for k in KEYS))
    tot.update(c)
print(f"{'ALL KERNELS':88s} {tot['total']:7d} " + " ".join(f"{tot[k]:7d}" for k in KEYS))
