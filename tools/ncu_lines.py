"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump per CUDA source line.
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > f.csv
    python tools/ncu_lines.py f.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] != "":   # a CUDA source line row (aggregated over its SASS)
        ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        try: out.append((int(r[ia]), int(r[isamp] or 0), fname, r[0], r[1].strip()))
        except ValueError: pass
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print("total warp instrs", tot, "samples", tots)
for n, s, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
    print(f"{n:11d} {100*n/tot:5.1f}% | smp {100*s/max(tots,1):5.1f}% | {f}:{ln:>4s} | {src[:105]}")
