#!/bin/bash
OUT=gpurun_out/r02ah
mkdir -p $OUT
i=0
for V in "WM_HOSTPIPE_ROWS=16" "WM_HOSTPIPE_ROWS=16 WM_HOSTPIPE_EVFLAGS=2" "WM_HOSTPIPE_ROWS=8" "WM_HOSTPIPE_ROWS=8 WM_HOSTPIPE_EVFLAGS=2"; do
  ( env $V WM_HOSTPIPE_TIME=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 6 --e2e-interval 1 2> $OUT/err_$i.txt | tail -1 ) > $OUT/bench_$i.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$i.json")); e = d["e2e"]
    print("%-44s e2e %.1f ms/step  chunks %s  ok=%s" % ("$V", e.get("ms_per_step", -1), e.get("host_pipe_chunks"), d["check"]["ok"]), e.get("error"))
except Exception as ex: print("$V", "ERR", ex)
PY
  grep "hostpipe" $OUT/err_$i.txt | tr '\n' ' '; echo
  i=$((i+1))
done
