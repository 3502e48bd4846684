#!/bin/bash
# multi-GPU visit: parity of the ring path against the oracle's N-slab world, then the weak-scaling bench line
# usage (under gpurun --gpus N): bash scripts/gpu_multi.sh <tag> <N>
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 ) > $OUT/pytest_multi.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e 2> $OUT/bench.err | tail -1 ) > $OUT/bench_n$N.json
( WM_CG=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_n${N}_cg0.json
cat $OUT/pytest_multi.log; tail -5 $OUT/bench.err
python - <<PY
import json
for f in ("bench_n$N.json", "bench_n${N}_cg0.json"):
    try:
        d = json.load(open("$OUT/" + f)); print(f, d["ms_per_step"], d["stage_ms"], d["run_info"]["cg_path"][:40], d["check"]["ok"], d["check"]["gauss_rel"])
    except Exception as e: print(f, "ERR", e)
PY
