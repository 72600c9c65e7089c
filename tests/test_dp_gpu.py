"""-m gpu, needs >= 2 GPUs (skipped otherwise): the data-parallel step on 2 ranks against the single-device
step on the same global batch (tools/dp_check.py under torchrun, NCCL)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("prec,B,mode", [("fp32", 16, "eager"), ("bf16", 64, "eager"), ("fp32", 16, "graph"),
                                         ("bf16", 64, "graph")])
def test_two_rank_step_equals_single_device(cuda, prec, B, mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py"), str(B), prec, mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "[dp_check] PASS" in r.stdout
