"""-m gpu: runs tests/multi_gpu_check.py under torchrun when the box has at least two B200s
(the single-GPU boxes skip it; the same host logic is covered on CPU by the gloo tests)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_collection_on_real_gpus(gpu):
    n = min(gpu.device_count(), 4)
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", "29631", os.path.join(root, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert f"multi_gpu_check ok on {n} GPUs" in res.stdout
