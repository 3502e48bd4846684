"""N-GPU == oracle N-slab (== 1 slab, tests/test_cpu.py) on real GPUs: NCCL ring exchange of ghost
rows, migrating particles, CG halos and all-reduces.  -m gpu; skipped with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("env", [{}, {"WM_INPLACE": "0"}])
def test_two_or_more_gpus_match_oracle(env):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    n = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(HERE, "mgpu_worker.py"), "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == n
