#!/bin/bash
# In-tile tail placement (k_fused_sm<TAIL> + k_place_rim) against its variants: parity tests + resident bench.
#   WM_SM=3: every cell changer left to k_place; WM_RIMPLACE=0: tail + the general k_place
# Usage: bash scripts/gpu_tail.sh <tag> [VAR=val ...]
TAG=${1:-tail}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/pytest_gpu.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
for v in ${@:-WM_SM=3}; do
( env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench_$v.err | tail -1 ) > $OUT/bench_$v.json
done
cat $OUT/pytest_gpu.log; python -c "
import json,glob
for f in sorted(glob.glob('$OUT/bench*.json')):
    j=json.load(open(f)); print(f, j['ms_per_step'], j['stage_ms'], j['config'].get('layout_rebuilds'))"
tail -3 $OUT/bench.err
