"""Launches tests/multigpu_check.py under torchrun when the box has >= 2 GPUs (skipped on 1-GPU boxes)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
def test_two_gpu_slabs_match_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(out.stdout[-4000:])
    print(out.stderr[-4000:])
    assert out.returncode == 0 and "MULTIGPU CHECK PASSED" in out.stdout
