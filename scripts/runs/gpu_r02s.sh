#!/bin/bash
OUT=gpurun_out/r02s
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multi.py::test_config1_sample_run_on_one_gpu 2>&1 | tail -8 ) > $OUT/pytest_gpu.log
cat $OUT/pytest_gpu.log
python scripts/run_configs.py shock --steps 200 --warmup 3 2>&1 | tail -1 > $OUT/shock_n1_200.json; cut -c1-700 $OUT/shock_n1_200.json
( timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2> $OUT/bench.err | tail -1 ) > $OUT/bench.json
python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['ms_per_step'],d['stage_ms'],d['check']['ok'])"
