#!/bin/bash
# final 1-GPU validation of the round-2 tree: the whole -m gpu suite, smoke(), the default bench (both arms)
OUT=gpurun_out/r02an
mkdir -p $OUT
timeout 2400 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1
tail -4 $OUT/pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
( timeout 900 python bench.py 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
python - <<PY
import json
d = json.load(open("$OUT/bench.json")); e = d["e2e"]
print("step %.3f ms  value %.3e  e2e %.1f ms/step (%.3e)  interval50 %.2f  roofline %.3f  whole %.3f  launches %s ok=%s" % (d["ms_per_step"], d["value"], e.get("ms_per_step", -1), e.get("value"), e.get("sync_interval_50", {}).get("ms_per_step", -1), d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], d["gpu_launches"], d["check"]["ok"]))
print("clocks", d["clocks"])
PY
tail -3 $OUT/bench.err
