#!/bin/bash
OUT=gpurun_out/r02v
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py::test_config1_sample_run_on_one_gpu 2>&1 | tail -8 ) > $OUT/pytest_gpu.log
cat $OUT/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused_dp" -s 3 -c 1 -o $OUT/prof_dp -f \
    python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > $OUT/ncu_full.log 2>&1
WM_SM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused_sm" -s 3 -c 1 -o $OUT/prof_sm -f \
    python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu > $OUT/ncu_full_sm.log 2>&1
for V in "WM_SM=5" "WM_SM=1"; do
( env $V timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>> $OUT/bench.err | tail -1 ) > $OUT/bench_$V.json
python -c "
import json;d=json.load(open('$OUT/bench_$V.json'));print('$V',d['ms_per_step'],d['stage_ms'],d['check']['ok'])"
done
