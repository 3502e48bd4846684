#!/bin/bash
# In-tile tail placement (k_fused_sm<TAIL>) against WM_SM=3 (every cell changer left to k_place): parity tests + resident bench.
# Usage: bash scripts/gpu_tail.sh <tag>
TAG=${1:-tail}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/pytest_gpu.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
( WM_SM=3 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench_sm3.err | tail -1 ) > $OUT/bench_sm3.json
cat $OUT/pytest_gpu.log; python -c "
import json
for f in ('bench','bench_sm3'):
    j=json.load(open('$OUT/%s.json'%f)); print(f, j['ms_per_step'], j['stage_ms'], j['config'].get('layout_rebuilds'))"
tail -3 $OUT/bench.err
