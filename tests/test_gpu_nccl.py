"""One process per GPU over NCCL (the bench.py / torchrun layout).  Needs >= 2 GPUs; skipped on a one-GPU box,
where tests/test_gpu_multirank.py covers the same schedule with virtual ranks."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("transport", ["nccl", "ipc"])
def test_two_processes_bit_identical(transport):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "medium", "3", transport]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "MULTI_GPU_CHECK PASS" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
