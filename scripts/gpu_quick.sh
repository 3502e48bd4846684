#!/bin/bash
# Quick GPU visit: parity tests, resident bench, ncu full capture of the particle kernels (small slab).
# Usage: bash scripts/gpu_quick.sh <tag> [noncu]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/pytest_gpu.log
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
if [ "$2" != "noncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused|k_place" -s 4 -c 2 -o $OUT/prof_pass \
    python bench.py --rows 128 --steps 2 --warmup 2 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
fi
cat $OUT/pytest_gpu.log $OUT/bench.json; tail -3 $OUT/bench.err
