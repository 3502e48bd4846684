#!/bin/bash
OUT=gpurun_out/r02af
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_hostpipe.py -q -m gpu > $OUT/pytest_hostpipe.log 2>&1
tail -5 $OUT/pytest_hostpipe.log | cut -c1-250
( WM_HOSTPIPE_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 2 --e2e-interval 1 2> $OUT/trace.err | tail -1 ) > $OUT/bench_trace.json
grep "chunks of" $OUT/trace.err
awk '/chunks of/{n++} n==3' $OUT/trace.err | grep -E "up +[0-5]:|up +3[01]:|dn +[0-3]:|dn +(2[89]|3[0-9]):"
i=0
for V in "WM_HOSTPIPE_ROWS=16" "WM_HOSTPIPE_ROWS=8" "WM_HOSTPIPE_ROWS=32"; do
  ( env $V timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 4 --e2e-interval 1 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_$i.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$i.json")); e = d["e2e"]
    print("%-22s step %.3f ms  e2e %.1f ms/step  chunks %s  ok=%s" % ("$V", d["ms_per_step"], e.get("ms_per_step", -1), e.get("host_pipe_chunks"), d["check"]["ok"]), e.get("error"))
except Exception as ex: print("$V", "ERR", ex)
PY
  i=$((i+1))
done
tail -3 $OUT/bench.err
