#!/bin/bash
# ncu --set full capture of one kernel under an env-selected variant: usage gpu_ncu.sh <tag> "VAR=1" <kernel regex> [rows]
TAG=$1; V=$2; RX=$3; ROWS=${4:-128}
OUT=gpurun_out/$TAG
mkdir -p $OUT
env $V timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 2 -c 1 -o $OUT/prof_pass -f \
    python bench.py --rows $ROWS --steps 2 --warmup 2 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
tail -2 $OUT/ncu_full.log
