#!/bin/bash
OUT=gpurun_out/r02as
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "deposit or host_step or dense or benchmark_strip" > $OUT/pytest.log 2>&1
tail -1 $OUT/pytest.log | cut -c1-200
for i in 0 1; do
  ( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_$i.json
  python - <<PY
import json
d = json.load(open("$OUT/bench_$i.json")); print("opaque gsh: step %.3f ms  fused %.3f  ok=%s" % (d["ms_per_step"], d["stage_ms"]["fused_push_deposit_boundary_sort"], d["check"]["ok"]))
PY
done
