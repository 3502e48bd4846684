#!/bin/bash
# pipelined wm_host_step: tests, then the bench with e2e (pipelined and WM_HOSTPIPE=0), chunk-size sweep
OUT=gpurun_out/r02ac
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_hostpipe.py -x -q -m gpu > $OUT/pytest_hostpipe.log 2>&1
tail -15 $OUT/pytest_hostpipe.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_step or push or deposit" > $OUT/pytest_parity_sub.log 2>&1
tail -3 $OUT/pytest_parity_sub.log
i=0
for V in "WM_HOSTPIPE=1" "WM_HOSTPIPE=0" "WM_HOSTPIPE_ROWS=8" "WM_HOSTPIPE_ROWS=32"; do
  ( env $V timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 4 --e2e-interval 1 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_$i.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$i.json")); e = d["e2e"]
    print("%-22s step %.3f ms (fused %.3f)  e2e %.1f ms/step  chunks %s  ok=%s" % ("$V", d["ms_per_step"], d["stage_ms"]["fused_push_deposit_boundary_sort"], e.get("ms_per_step", -1), e.get("host_pipe_chunks"), d["check"]["ok"]), e.get("error"))
except Exception as ex: print("$V", "ERR", ex)
PY
  i=$((i+1))
done
tail -3 $OUT/bench.err
