#!/bin/bash
# 8-GPU visit: ring parity at 8 ranks, config 1 on 4 GPUs, weak-scaling bench lines at N = 4, 8 (+ the host-loop CG for
# comparison), the non-uniform configs at 8 GPUs, two points of the config-5 sweep at 8 GPUs.
TAG=${1:-r02w_n8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "two_or_more or four_gpus" 2>&1 | tail -30 ) > $OUT/pytest_multi.log
cat $OUT/pytest_multi.log | tail -5
for N in 8 4; do
  ( timeout 600 $TR --nproc-per-node $N --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 --no-e2e 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_n$N.json
done
( WM_CG=0 WM_MIGSYNC=1 timeout 600 $TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_n8_r01path.json
( timeout 600 $TR --nproc-per-node 8 --master-port 29622 bench.py --gpus 8 --steps 10 --warmup 3 --e2e-steps 2 --e2e-interval 0 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_n8_e2e.json
( timeout 900 $TR --nproc-per-node 8 --master-port 29623 scripts/run_configs.py harris --gpus 8 --steps 20 --warmup 3 2>> $OUT/bench.err | tail -1 ) > $OUT/harris_n8.json
( timeout 900 $TR --nproc-per-node 8 --master-port 29624 scripts/run_configs.py shock --gpus 8 --steps 200 --warmup 3 2>> $OUT/bench.err | tail -1 ) > $OUT/shock_n8.json
( timeout 600 $TR --nproc-per-node 8 --master-port 29625 bench.py --gpus 8 --nx 8192 --rows 1024 --ppc 16 --steps 10 --warmup 3 --no-e2e 2>> $OUT/bench.err | tail -1 ) > $OUT/sweep_8192sq_p16_n8.json
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/*.json")):
    try:
        d = json.load(open(f))
        ms = d.get("ms_per_step", d.get("ms_per_step_device"))
        print(f.split("/")[-1], "ms/step", ms, "G p-steps/s", round(d.get("value", d.get("particle_steps_per_s", 0)) / 1e9, 2), d.get("check", {}).get("ok"), d.get("stage_ms"), (d.get("e2e") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 $OUT/bench.err
