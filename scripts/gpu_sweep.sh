#!/bin/bash
# BASELINE configs[4]: push+deposit microbenchmark sweep on one GPU (uniform plasma, ppc x grid); one JSON line per point
OUT=gpurun_out/sweep; mkdir -p $OUT
: > $OUT/sweep.jsonl
for P in "1024 1024 16" "1024 1024 64" "1024 1024 256" "2048 2048 16" "2048 2048 64" "4096 512 16" "4096 512 128" "8192 256 64"; do
  set -- $P
  timeout 300 python bench.py --nx $1 --rows $2 --ppc $3 --steps 8 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 >> $OUT/sweep.jsonl
done
python - <<'PY'
import json
print("| grid | ppc/species | particles | ms/step | G particle-steps/s | % of 192 B roofline | fused ms | field ms | place ms | CG iters |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for ln in open("gpurun_out/sweep/sweep.jsonl"):
    ln = ln.strip()
    if not ln.startswith("{"): continue
    j = json.loads(ln); s = j["stage_ms"]; w = j["config"]["workload"]
    grid = w.split()[1]; ppc = w.split("grid")[1].split(",")[1].strip().split()[0]
    print("| %s | %s | %d | %.2f | %.2f | %.1f | %.2f | %.2f | %.2f | %s |" % (grid, ppc, j["config"]["particles"], j["ms_per_step"], j["value"] / 1e9,
          100 * j["roofline"]["whole_step"]["frac"], s["fused_push_deposit_boundary_sort"], s["field_solve"], s["sort_place_cell_changers"], j["config"]["cg_iters"]))
PY
