"""Row-sharded scene over real NCCL ranks (one process per GPU).  Skipped unless >= 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("k", [1, 4])
def test_row_sharded_scene_over_nccl(k):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + k), os.path.join(ROOT, "tests", "sharded_nccl_worker.py"), str(k)]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "OK" in res.stdout
