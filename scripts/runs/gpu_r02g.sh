#!/bin/bash
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 ) > $OUT/pytest_gpu.log
( WM_CGTRACE=20 timeout 600 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | grep cgtrace | head -16 ) > $OUT/trace.txt
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_launch_bench.log 2>&1
cat $OUT/pytest_gpu.log $OUT/trace.txt; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['ms_per_step'],d['stage_ms'],d['check']['ok'])"
grep k_cg_persist $OUT/launches.csv | tail -1
