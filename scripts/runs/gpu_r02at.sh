#!/bin/bash
# ptxas register-usage-level variants of k_fused_sm's file (scripts/build_variants.sh), A/B on the benchmark slab
OUT=gpurun_out/r02at
mkdir -p $OUT
for V in default r0 r7 r10 default r0; do
  L=""; [ $V != default ] && L="WM_LIB=$PWD/wumingpic2d_b200/variants/lib_$V.so"
  ( env $L timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_$V.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$V.json")); print("%-8s step %.3f ms  fused %.3f  ok=%s" % ("$V", d["ms_per_step"], d["stage_ms"]["fused_push_deposit_boundary_sort"], d["check"]["ok"]))
except Exception as e: print("$V", "ERR", e)
PY
done
tail -2 $OUT/bench.err
