"""Launches tests/multi_gpu_check.py under torchrun when the box has at least two GPUs (the driver's 1-GPU
test run skips it; `gpurun --gpus 2` and the scaling run exercise it)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ray_sharded_forward_pass_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = str(29600 + os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", port, os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    if r.returncode != 0 or "MULTI_GPU_CHECK ok" not in r.stdout:
        k = r.stdout.find("MULTI_GPU_CHECK failed")
        pytest.fail(r.stdout[k:k + 4000] if k >= 0 else r.stdout[-4000:])
