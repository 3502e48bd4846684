#!/bin/bash
# N GPUs: ring parity incl. the pipelined host step, then the bench line with e2e (pipelined and not)
N=${1:-2}
OUT=gpurun_out/r02aj_n$N
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "two_or_more" 2>&1 | tail -30 ) > $OUT/pytest_multi.log
tail -5 $OUT/pytest_multi.log | cut -c1-300
i=0
for V in "WM_HOSTPIPE=1" "WM_HOSTPIPE=0"; do
  ( env $V timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$i bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 3 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_$i.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$i.json")); e = d["e2e"]
    print("%-16s N=%d step %.3f ms  e2e %.1f ms/step  chunks %s  interval50 %.2f  ok=%s" % ("$V", d["n_gpus"], d["ms_per_step"], e.get("ms_per_step", -1), e.get("host_pipe_chunks"), e.get("sync_interval_50", {}).get("ms_per_step", -1), d["check"]["ok"]), e.get("error"), e.get("cpu_affinity_rank0"))
except Exception as ex: print("$V", "ERR", ex)
PY
  i=$((i+1))
done
tail -3 $OUT/bench.err
