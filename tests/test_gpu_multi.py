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
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(HERE, "mgpu_worker.py"), "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == n


def _run_config1(n, steps):
    cmd = [sys.executable, os.path.join(HERE, "config1_worker.py"), str(steps)]
    if n > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
               "--master-port", "29549"] + cmd[1:]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1700)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" ok") == n
    return r.stdout


def test_config1_sample_run_on_four_gpus():
    """BASELINE configs[0] exactly: proj/weibel/config_sample.json -- 256 x 256, 20 ppc, num_process = 4, 1000 steps -- on four
    GPUs (one 64-row slab each, ring migration, in-kernel CG exchange) against the oracle's 4-slab world: one-step parity, then
    the energy history every 50 steps within 1e-6 relative (tests/config1_worker.py)."""
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs (gpurun --gpus 4)")
    _run_config1(4, 1000)


def test_config1_sample_run_on_one_gpu():
    """The same run on ONE GPU holding all 256 rows, against the oracle's 4-slab world: the decomposition must not matter."""
    _run_config1(1, 1000)
