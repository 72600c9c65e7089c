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


def test_run_py_under_torchrun(cuda):
    """``torchrun -m eadgan_b200.run script.py`` must all-reduce gradients although the unmodified script
    never mentions data parallelism (every Adam built after parallel.init_from_env() attaches itself).  Two ranks draw
    different data; after 4 steps their parameters and BatchNorm running statistics must be bit-identical."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = os.path.join(ROOT, "tests", "scripts", "dp_mini.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", "-m", "eadgan_b200.run", script]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), EADGAN_PRECISION="fp32")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    rows = sorted((json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")), key=lambda d: d["rank"])
    assert len(rows) == 2 and rows[0]["opt"] == "eadgan_b200.optim"
    assert rows[0]["checksum"] == rows[1]["checksum"]        # replicas identical (weights + BN running statistics)
    assert rows[0]["loss"] != rows[1]["loss"]                # ... although they saw different data
