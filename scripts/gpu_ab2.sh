#!/bin/bash
# A/B of environment variants on the benchmark slab: usage gpu_ab2.sh <tag> "VAR=a" "VAR=b OTHER=c" ...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
i=0
for V in "$@"; do
  ( env $V timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_$i.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$i.json")); print("%-28s step %.3f ms  %s ok=%s" % ("$V", d["ms_per_step"], {k: round(v, 3) for k, v in d["stage_ms"].items()}, d["check"]["ok"]))
except Exception as e: print("$V", "ERR", e)
PY
  i=$((i+1))
done
tail -3 $OUT/bench.err
