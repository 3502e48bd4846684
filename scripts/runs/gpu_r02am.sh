#!/bin/bash
OUT=gpurun_out/r02am_n8
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
lscpu | grep -iE "numa|model name|socket|^cpu\(s\)" > $OUT/lscpu.txt 2>&1
free -g > $OUT/free.txt 2>&1
( timeout 400 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29631 scripts/micro/pcie_ranks.py 2 2> $OUT/err8.txt | tail -1 ) > $OUT/pcie_n8.json
( timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 4 --master-port 29632 scripts/micro/pcie_ranks.py 2 2> $OUT/err4.txt | tail -1 ) > $OUT/pcie_n4.json
python - <<PY
import json
for f in ("pcie_n8.json", "pcie_n4.json"):
    try:
        d = json.load(open("$OUT/" + f))
        print(f, d["numa_nodes_online"], d["host_cpus"], json.dumps({k: {m: round(v, 1) for m, v in t.items()} for k, t in d["sum_GBps_per_direction"].items()}))
        print("  cpus", [r["cpus"] for r in d["ranks"]])
        print("  both after binding per rank", [round(r["rates_GBps"]["hostalloc_after_binding"]["both"], 1) for r in d["ranks"]])
    except Exception as ex: print(f, "ERR", ex)
PY
cat $OUT/lscpu.txt; head -12 $OUT/topo.txt | cut -c1-160; tail -3 $OUT/err8.txt
