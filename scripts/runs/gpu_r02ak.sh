#!/bin/bash
# layout rebuild, count-driven copy: parity of the overflow paths, then the shock run (config 4 slab) with the rebuild phases timed
OUT=gpurun_out/r02ak
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sources.py tests/test_gpu_hostpipe.py -x -q -m gpu -k "sort_variants or sources or shock or overflow or rebuild" > $OUT/pytest.log 2>&1
tail -3 $OUT/pytest.log | cut -c1-300
( WM_REBUILD_TIME=1 timeout 900 python scripts/run_configs.py shock --steps 200 --warmup 3 2> $OUT/shock.err | tail -1 ) > $OUT/shock.json
python - <<PY
import json
d = json.load(open("$OUT/shock.json"))
print("shock: %.2f ms/step  %.2f G/s  rebuilds %d  stages %s" % (d["ms_per_step_device"], d["particle_steps_per_s"] / 1e9, d["layout_rebuilds_in_timed_steps"], {k: round(v, 2) for k, v in d["stage_ms"].items()}))
PY
grep rebuild $OUT/shock.err | head -5; grep rebuild $OUT/shock.err | tail -3
grep -c rebuild $OUT/shock.err
