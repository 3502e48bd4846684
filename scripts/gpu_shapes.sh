#!/bin/bash
# same particle count, different grid aspect: row stride in the particle store changes (TLB reach test)
mkdir -p gpurun_out/shapes
for S in "4096 512" "1024 2048" "256 8192" "16384 128"; do
  set -- $S
  ( timeout 600 python bench.py --nx $1 --rows $2 --steps 6 --warmup 3 --no-e2e --no-cpu 2> gpurun_out/shapes/err_$1.txt | tail -1 ) > gpurun_out/shapes/b_$1.json
  python -c "
import json
j=json.load(open('gpurun_out/shapes/b_$1.json')); print('nx $1 rows $2', j['ms_per_step'], j['stage_ms'])" || tail -3 gpurun_out/shapes/err_$1.txt
done
