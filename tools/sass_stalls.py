"""Static issue-cycle model of a kernel's SASS: every instruction carries the number of cycles the warp must wait before its next
instruction may issue (control bits 105..108 = bits 41..44 of the second 64-bit word).  Summing them over a loop body gives the
single-warp issue time of one iteration when nothing but the compiler-scheduled fixed latencies holds the warp back.
    cuobjdump -sass X.o > x.sass ; python tools/sass_stalls.py x.sass <kernel-substring> [lo_addr hi_addr]
Without an address range: prints every backward branch (loop) with the instruction count and stall sum of its body."""
import re, sys
txt = open(sys.argv[1]).read().split("Function :")
pat = sys.argv[2]
body = [t for t in txt if pat in t.split("\n")[0]][0]
ins = []
lines = body.split("\n")
i = 0
while i < len(lines):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", lines[i])
    if m and i + 1 < len(lines):
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        hi = int(m2.group(1), 16) if m2 else 0
        stall = (hi >> 41) & 0xF
        yld = (hi >> 45) & 1
        ins.append((int(m.group(1), 16), m.group(2).strip(), stall, yld))
        i += 2
    else:
        i += 1
print(f"{len(ins)} instructions, stall sum {sum(s for _, _, s, _ in ins)}")
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    sel = [x for x in ins if lo <= x[0] <= hi]
    print(f"[{lo:#x}, {hi:#x}]: {len(sel)} instructions, stall sum {sum(s for _, _, s, _ in sel)}")
    import collections
    c = collections.Counter()
    for a, t, s, y in sel:
        op = re.sub(r"^@!?U?P[0-9T]+\s+", "", t).split()[0].split(".")[0]
        c[op] += s
    print("stall cycles by opcode:", c.most_common(14))
else:
    for a, t, s, y in ins:
        m = re.search(r"BRA(?:\.U)?(?:\.ANY)?\s+(?:!?U?P[0-9T]+,\s+)?(0x[0-9a-f]+)", t)
        if m and int(m.group(1), 16) <= a:
            lo = int(m.group(1), 16)
            sel = [x for x in ins if lo <= x[0] <= a]
            print(f"loop [{lo:#x}, {a:#x}]: {len(sel)} instructions, stall sum {sum(s for _, _, s, _ in sel)}")
