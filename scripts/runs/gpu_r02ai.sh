#!/bin/bash
# round-2 validation after the pipelined host step: sanitizer on it, the whole -m gpu suite, the default bench
OUT=gpurun_out/r02ai
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
for T in memcheck racecheck; do
  ( timeout 900 $CS --tool $T --print-limit 20 python scripts/sanitize_driver.py hostpipe 2 2>&1 | tail -40 ) > $OUT/${T}_hostpipe.log
  echo "== ${T}_hostpipe: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver' $OUT/${T}_hostpipe.log | tr '\n' ' ')"
done
timeout 2400 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1
tail -6 $OUT/pytest_gpu.log | cut -c1-300
( timeout 900 python bench.py 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
python - <<PY
import json
d = json.load(open("$OUT/bench.json")); e = d["e2e"]
print("step %.3f ms  value %.3e  e2e %.1f ms/step (%.3e)  interval50 %.2f ms/step  roofline %.3f  whole %.3f  cpu %s  ok=%s" % (d["ms_per_step"], d["value"], e.get("ms_per_step", -1), e.get("value"), e.get("sync_interval_50", {}).get("ms_per_step", -1), d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], d["cpu_baseline"], d["check"]["ok"]))
PY
tail -3 $OUT/bench.err
