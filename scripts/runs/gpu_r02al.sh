#!/bin/bash
# final 8-GPU visit of round 2: ring parity at 8 ranks incl. the pipelined host step, weak-scaling lines with e2e at N = 8, 4,
# e2e without the pipeline for comparison, two more points of the config-5 sweep (4096^2 global grid, 16 and 128 ppc)
OUT=gpurun_out/r02al_n8
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "two_or_more and env0" 2>&1 | tail -30 ) > $OUT/pytest_multi.log
tail -3 $OUT/pytest_multi.log | cut -c1-300
( timeout 900 $TR --nproc-per-node 8 --master-port 29618 bench.py --gpus 8 --steps 20 --warmup 5 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_n8.json
( timeout 900 $TR --nproc-per-node 4 --master-port 29614 bench.py --gpus 4 --steps 20 --warmup 5 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_n4.json
( WM_HOSTPIPE=0 timeout 900 $TR --nproc-per-node 8 --master-port 29628 bench.py --gpus 8 --steps 5 --warmup 3 --e2e-steps 2 --e2e-interval 0 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_n8_nopipe.json
( timeout 600 $TR --nproc-per-node 8 --master-port 29625 bench.py --gpus 8 --ppc 16 --steps 10 --warmup 3 --no-e2e 2>> $OUT/bench.err | tail -1 ) > $OUT/sweep_4096sq_p16_n8.json
( timeout 600 $TR --nproc-per-node 8 --master-port 29626 bench.py --gpus 8 --ppc 128 --steps 10 --warmup 3 --no-e2e 2>> $OUT/bench.err | tail -1 ) > $OUT/sweep_4096sq_p128_n8.json
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/*.json")):
    try:
        d = json.load(open(f))
        e = d.get("e2e") or {}
        print(f.split("/")[-1], "N", d.get("n_gpus"), "ms/step", round(d["ms_per_step"], 3), "G p-steps/s", round(d["value"] / 1e9, 2), "ok", d["check"]["ok"], {k: round(v, 2) for k, v in d["stage_ms"].items()},
              "e2e", e.get("ms_per_step"), "chunks", e.get("host_pipe_chunks"), "i50", (e.get("sync_interval_50") or {}).get("ms_per_step"), e.get("error"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -4 $OUT/bench.err
