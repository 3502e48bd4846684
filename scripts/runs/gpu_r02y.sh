#!/bin/bash
OUT=gpurun_out/r02y
mkdir -p $OUT
( timeout 1000 python -m pytest tests/test_gpu_loopback.py tests/test_gpu_parity.py::test_host_steps_sync_interval tests/test_gpu_parity.py::test_benchmark_strip_matches_oracle tests/test_snapshot.py -m gpu -x -q 2>&1 | tail -30 ) > $OUT/pytest.log
cat $OUT/pytest.log
for W in 0 3 20 60 150; do
  ( timeout 600 python bench.py --steps 3 --warmup $W --no-e2e --no-cpu 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_w$W.json
  python -c "
import json;d=json.load(open('$OUT/bench_w$W.json'));print('warmup $W', d['ms_per_step'], d['stage_ms']['fused_push_deposit_boundary_sort'], d['stage_ms']['field_solve'])"
done
