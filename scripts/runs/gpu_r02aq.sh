#!/bin/bash
# final evidence of round 2 on the final tree: launch list of the bench command + one --set full capture of k_fused_sm
OUT=gpurun_out/r02aq
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > $OUT/ncu_launch_bench.log 2>&1
tail -1 $OUT/ncu_launch_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fused_sm" -s 3 -c 1 -o $OUT/prof_sm -f \
    python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > $OUT/ncu_full_sm.log 2>&1
tail -2 $OUT/ncu_full_sm.log | cut -c1-200
( timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
ls -la $OUT
