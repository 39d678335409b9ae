"""Static SASS opcode counts per kernel of the shipped library -> profiles/sass_counts_rNN.md
    python tools/sass_counts.py [lib.so] > profiles/sass_counts_r02.md"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "medical_image_classification_b200/lib/libb200ssm.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "ELECT", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "MUFU", "HMMA",
       "SHFL", "LDGSTS", "LDS", "STS", "REDG", "ATOMG", "LDL", "STL"]
per, name = collections.OrderedDict(), None
for ln in txt.split("\n"):
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", ln)
    if m and name:
        per[name]["instrs"] += 1
        per[name][m.group(1)] += 1
dem = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.split("\n")
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("# SASS opcode counts of the shipped libb200ssm.so (round 2, final build)\n")
print("`python tools/sass_counts.py` = `cuobjdump -sass medical_image_classification_b200/lib/libb200ssm.so`, static counts per kernel (whole kernel, not the hot "
      "loop).  `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` / `UTMAREDG` = TMA tensor load / store / reduce-add, `UBLKCP` = 1-D bulk "
      "copy, `ELECT` = elect.sync (single-thread issue), `SYNCS` = mbarrier ops, `FFMA2` / `FMUL2` / `FADD2` = packed f32x2, `HMMA` = mma.sync (legacy tensor "
      "path), `LDGSTS` = cp.async, `LDL` / `STL` = local-memory spills.\n")
print("Library totals: " + ", ".join(f"{o} {tot[o]}" for o in ops) + "\n")
print("| kernel | instrs | " + " | ".join(ops) + " |")
print("|---|---:|" + "---:|" * len(ops))
for (n, c), d in zip(per.items(), dem):
    d = re.sub(r"\(.*", "", d)
    print(f"| `{d[:100]}` | {c['instrs']} | " + " | ".join(str(c[o]) if c[o] else "" for o in ops) + " |")
