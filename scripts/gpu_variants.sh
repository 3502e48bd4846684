#!/bin/bash
# Bench several env-selected variants back to back: usage gpu_variants.sh <tag> "VAR=1" "VAR=2" ...
# The first variant also runs the GPU parity tests.
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
i=0
for V in "$@"; do
  N=$(echo "$V" | tr ' =' '__')
  if [ $i -eq 0 ]; then ( env $V timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_$N.log; cat $OUT/pytest_$N.log; fi
  ( env $V timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench_$N.err | tail -1 ) > $OUT/bench_$N.json
  python -c "
import json
j=json.load(open('$OUT/bench_$N.json')); print('$V', j['ms_per_step'], j['stage_ms'])" || tail -5 $OUT/bench_$N.err
  i=$((i+1))
done
