"""world_size-2 tests on CPU (torch.distributed, gloo): the host-side logic of the multi-GPU path and the reference arm
of bench.py under torchrun (rank 0 alone runs and prints, the other rank exits 0)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _torchrun(args, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_slab_logic_and_harness_reductions_two_ranks():
    r = _torchrun([os.path.join(HERE, "gloo_worker.py")], 29561)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count(" ok") == 2


def test_bench_reference_arm_under_torchrun():
    r = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], 29562)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "rank 0 alone prints the line"
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["value"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["cpu_baseline"]["kind"] == "port"
