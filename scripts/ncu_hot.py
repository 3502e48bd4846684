#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel in an ncu report:  ncu_hot.py report.ncu-rep kernel_regex [N]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hs = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hs[0]]; ci = {n: i for i, n in enumerate(h)}
end = hs[1] - 1 if len(hs) > 1 else len(rows)
body = [r for r in rows[hs[0] + 1:end] if len(r) >= len(h) and r[0] != 'Address']
tot = sum(int(r[ci['# Samples']]) for r in body)
print('total samples', tot, 'instructions', len(body))
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg = {s: sum(int(r[ci[s]]) for r in body) for s in stalls}
print(sorted(((v, k) for k, v in agg.items()), reverse=True)[:6])
for r in sorted(body, key=lambda r: -int(r[ci['# Samples']]))[:N]:
    st = sorted([(int(r[ci[s]]), s) for s in stalls], reverse=True)[:2]
    print(r[ci['# Samples']], r[ci['Instructions Executed']], r[ci['Source']][:64], st)
