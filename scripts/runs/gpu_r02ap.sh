#!/bin/bash
# after factoring the pipelined step's schedule into hp_schedule: parity tests + e2e
OUT=gpurun_out/r02ap
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_hostpipe.py tests/test_gpu_parity.py -x -q -m gpu -k "hostpipe or host_step or pipelined or register" > $OUT/pytest.log 2>&1
tail -3 $OUT/pytest.log | cut -c1-300
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 4 --e2e-interval 1 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
python - <<PY
import json
d = json.load(open("$OUT/bench.json")); e = d["e2e"]
print("step %.3f ms  e2e %.1f ms/step chunks %s ok=%s" % (d["ms_per_step"], e.get("ms_per_step", -1), e.get("host_pipe_chunks"), d["check"]["ok"]), e.get("error"))
PY
tail -2 $OUT/bench.err
