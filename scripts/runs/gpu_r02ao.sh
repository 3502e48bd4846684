#!/bin/bash
# ncu launch list of the pipelined wm_host_step on the sanitizer's small world (six chunks): which kernels a call launches
OUT=gpurun_out/r02ao
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/hostpipe_launches.csv python scripts/sanitize_driver.py hostpipe 1 > $OUT/run.log 2>&1
tail -2 $OUT/run.log
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/hostpipe_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value")
c = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0]
    c.setdefault(k, [0, 0.0]); c[k][0] += 1; c[k][1] += float(r[vi].replace(",", ""))
for k, (n, t) in c.items(): print("%-40s %4d launches %10.1f us" % (k[:40], n, t / 1e3))
PY
