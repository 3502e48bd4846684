#!/bin/bash
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( WM_CGTRACE=20 timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | grep cgtrace ) > $OUT/trace.txt
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
cat $OUT/trace.txt; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['ms_per_step'],d['stage_ms'],d['check']['ok'])"
