"""Multi-GPU paths under torchrun (-m gpu; skipped on a single-GPU box): window sharding, row-sharded Gram + NCCL
all-reduce, and ONE ADMM problem sharded over the GPUs with device-initiated peer stores."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch

    return torch.cuda.device_count()


def _torchrun(script, n, port, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script), *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("script,port,arg", [("dist_check.py", 29561, None), ("admm_shard_check.py", 29562, None),
                                             ("admm_shard_check.py", 29563, "ball"), ("admm_shard_check.py", 29564, "lpv")])
def test_two_gpu_paths(script, port, arg):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _torchrun(script, 2, port, *([arg] if arg else []))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
