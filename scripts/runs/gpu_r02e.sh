#!/bin/bash
TAG=${1:-r02e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cg_persist" -s 2 -c 1 -o $OUT/prof_cg -f \
    python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
cat $OUT/pytest_gpu.log; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['ms_per_step'],d['stage_ms'],d['check'])"
grep k_cg_persist $OUT/launches.csv | tail -2
