#!/bin/bash
# A/B of one env-selected variant: usage gpu_ab.sh <tag> VAR=1 [kernel regex]
TAG=$1; export $2; RX=${3:-k_fused}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 2 -c 1 -o $OUT/prof_pass \
    python bench.py --rows 128 --steps 2 --warmup 2 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
cat $OUT/pytest_gpu.log; python -c "
import json,sys
j=json.load(open('$OUT/bench.json')); print(j['ms_per_step'], j['stage_ms'])"
tail -3 $OUT/bench.err
