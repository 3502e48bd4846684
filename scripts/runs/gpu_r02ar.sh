#!/bin/bash
# word 0 of the next particle loaded an iteration ahead (WM_SM=12): parity, then A/B on the benchmark slab
OUT=gpurun_out/r02ar
mkdir -p $OUT
WM_SM=12 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "deposit or host_step or dense or benchmark_strip or roundtrip" > $OUT/pytest_sm12.log 2>&1
tail -2 $OUT/pytest_sm12.log | cut -c1-200
i=0
for V in "WM_SM=1" "WM_SM=12" "WM_SM=1" "WM_SM=12"; do
  ( env $V timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_$i.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$i.json")); print("%-10s step %.3f ms  fused %.3f  ok=%s" % ("$V", d["ms_per_step"], d["stage_ms"]["fused_push_deposit_boundary_sort"], d["check"]["ok"]))
except Exception as e: print("$V", "ERR", e)
PY
  i=$((i+1))
done
tail -2 $OUT/bench.err
