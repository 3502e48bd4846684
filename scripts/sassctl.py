#!/usr/bin/env python
"""SASS with decoded control fields (stall, yield, write/read barrier, wait mask) for one kernel.
   usage: sassctl.py <object> <function substring> [lo_hex hi_hex]"""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
on = False
i = 0
while i < len(out):
    ln = out[i]
    if "Function :" in ln:
        on = pat in ln
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;\s*/\* 0x([0-9a-f]{16}) \*/", ln) if on else None
    if m and i + 1 < len(out):
        m2 = re.search(r"/\* 0x([0-9a-f]{16}) \*/", out[i + 1])
        a = int(m.group(1), 16)
        if m2 and lo <= a <= hi:
            hiw = int(m2.group(1), 16)
            ctl = hiw >> 41  # bits 105.. of the 128-bit word
            stall = ctl & 0xf
            yld = (ctl >> 4) & 1
            wb = (ctl >> 5) & 7
            rb = (ctl >> 8) & 7
            wm = (ctl >> 11) & 0x3f
            s = "%05x  st%-2d %s W%s R%s wait[%s]  %s" % (a, stall, "Y" if yld == 0 else " ", "-" if wb == 7 else wb, "-" if rb == 7 else rb,
                                                         "".join(str(b) if wm >> b & 1 else "." for b in range(6)), m.group(2))
            print(s)
        i += 2
        continue
    i += 1
