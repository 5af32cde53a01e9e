"""One process per GPU over NCCL (torchrun): N ranks, each with its own handle and its shard of the iterations, must
reproduce what one GPU computes -- the path bench.py's device-resident leg and scripts/bench_multi.py use.
Needs two GPUs (skipped below); the single-process device group has its own tests (test_gpu_group.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _run(n, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "scripts", "bench_multi.py"), "--config", "2", "--iters", "96"]
    if n == 1:
        cmd = [sys.executable, os.path.join(ROOT, "scripts", "bench_multi.py"), "--config", "2", "--iters", "96"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


@pytest.mark.skipif("_n_gpus() < 2")
def test_nccl_ranks_reproduce_one_gpu():
    one = _run(1, 29611)
    two = _run(2, 29612)
    assert two["n_gpus"] == 2 and two["ranks_agree"]
    assert two["checksum"] == one["checksum"]          # the gathered [R x G] matrix is the one-GPU matrix
    if _n_gpus() >= 4:
        four = _run(4, 29613)
        assert four["ranks_agree"] and four["checksum"] == one["checksum"]
