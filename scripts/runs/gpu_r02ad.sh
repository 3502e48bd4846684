#!/bin/bash
OUT=gpurun_out/r02ad
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_hostpipe.py -q -m gpu > $OUT/pytest_hostpipe.log 2>&1
tail -25 $OUT/pytest_hostpipe.log | cut -c1-250
( WM_HOSTPIPE_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 2 --e2e-interval 1 2> $OUT/trace.err | tail -1 ) > $OUT/bench_trace.json
python - <<PY
import json
d = json.load(open("$OUT/bench_trace.json")); e = d["e2e"]
print("e2e %.1f ms/step chunks %s" % (e.get("ms_per_step", -1), e.get("host_pipe_chunks")), e.get("error"))
PY
grep hostpipe $OUT/trace.err | tail -75
