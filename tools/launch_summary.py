"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel.
    python tools/launch_summary.py gpurun_out/launches_r01.csv [steps] > profiles/launches_r01_summary.md"""
import csv, re, sys
from collections import defaultdict
path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void ", "", name)
    ms = float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    a = agg[name]
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print(f"# ncu launch list summary: {path}\n")
print(f"{len(rows)-1} launches, {tot:.3f} ms total device time ({tot/steps:.3f} ms per step over {steps} steps); cold-cache, serialised: compare shares\n")
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{name[:110]}` | {n} | {ms:.3f} | {100*ms/tot:.1f}% |")
