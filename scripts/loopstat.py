#!/usr/bin/env python
"""Static instruction mix of the innermost loop that contains the first LDS.128 of a kernel.
   usage: loopstat.py <object> <function substring> [--list]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins, on = [], False
for ln in out.splitlines():
    if "Function :" in ln:
        on = pat in ln
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;\s*/\*", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
addr = [a for a, _ in ins]
first = next(a for a, t in ins if "LDS.128" in t)
best = None
for a, t in ins:
    m = re.search(r"BRA.*?(0x[0-9a-f]+)", t)
    if m and "ANY" not in t:
        tgt = int(m.group(1), 16)
        if tgt <= first <= a and (best is None or a - tgt < best[1] - best[0]):
            best = (tgt, a)
lo, hi = best
body = [(a, t) for a, t in ins if lo <= a <= hi]
def opc(t):
    p = t.split()
    op = p[1] if p[0].startswith("@") else p[0]
    return op.split(".")[0]
h = collections.Counter(opc(t) for _, t in body)
fp64 = sum(h[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
print("kernel %d instr; loop 0x%x..0x%x: %d instr, FP64 %d, other %d, issue-cycle estimate %d" % (len(ins), lo, hi, len(body), fp64, len(body) - fp64, 2 * fp64 + len(body) - fp64))
print("  ".join("%s %d" % kv for kv in h.most_common(28)))
if "--list" in sys.argv:
    for a, t in body:
        if opc(t) not in ("DFMA", "DMUL", "DADD", "LDS"):
            print("%05x %s" % (a, t))
