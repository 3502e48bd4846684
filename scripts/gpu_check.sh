#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list and a full capture of the two particle kernels.
# Usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
( timeout 900 python bench.py --steps 10 --warmup 3 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
( timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_launch_bench.log 2>&1
# full capture of the particle kernels at the benchmark size (one launch each, after two warm-up steps)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused_sm|k_place_rim" -s 4 -c 2 -o $OUT/prof_pass -f \
    python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_cg_pap|k_cg_update2" -s 12 -c 2 -o $OUT/prof_cg -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_cg.log 2>&1
ls -la $OUT
cat $OUT/pytest_gpu.log $OUT/bench.json $OUT/bench_ref.json
