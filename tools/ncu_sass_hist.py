"""Dynamic opcode histogram + stall totals from an ncu SASS source page.
    ncu -i X.ncu-rep --page source --csv --print-source sass --kernel-name regex:K > f.csv
    python tools/ncu_sass_hist.py f.csv [units]      # units: divide counts by this many (e.g. chunk executions)"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = rows[1]
ia, isrc, ismp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [(k, c) for k, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
ops = collections.Counter(); smp = collections.Counter(); stalls = collections.Counter(); tot = 0; tsmp = 0
for r in rows[2:]:
    if len(r) < len(h) or r[ia] == "Address": continue
    s = r[isrc].strip()
    s = re.sub(r"^@!?U?P[0-9T]+\s+", "", s)
    op = s.split()[0].rstrip(";")
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "MUFU", "SHFL", "LDG", "STG")) else op.split(".")[0]
    n = int(r[iex] or 0); m = int(r[ismp] or 0)
    ops[op] += n; smp[op] += m; tot += n; tsmp += m
    for k, c in stall_cols: stalls[c] += int(r[k] or 0)
print(f"total warp instructions {tot} ({tot / units:.1f} per unit), samples {tsmp}")
for op, n in ops.most_common(45):
    print(f"{op:14s} {n / units:10.1f} {100 * n / tot:5.1f}%   samples {100 * smp[op] / max(tsmp, 1):5.1f}%")
print("stall totals:")
for c, n in stalls.most_common(12): print(f"  {c:28s} {100 * n / max(tsmp, 1):5.1f}%")
