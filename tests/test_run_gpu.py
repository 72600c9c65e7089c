"""-m gpu: the drop-in claim at SCRIPT level.  tests/scripts/mini_infogan.py is a flat training script in the style
of the reference's (stock torch.nn names, spectral_norm imported from torch, F.sigmoid / F.softmax, BCE / CE / MSE,
two torch.optim.Adam, loop at import time).  It is run UNCHANGED (a) with stock PyTorch on the GPU and (b) under
``python -m eadgan_b200.run`` in fp32 and bf16 mode; seeded inputs and weights are identical, so the printed losses
must agree.  (The real reference scripts are exercised the same way in the build container, where /root/reference
exists: oracle/ref_runner.py; they need datasets that are not available offline.)"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "tests", "scripts", "mini_infogan.py")


def _run(cmd, env_extra):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), **env_extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    rows = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(rows) == 3, r.stdout
    return rows


def test_unmodified_script_under_the_shim_matches_stock_torch(cuda):
    args = ["--n_iter", "3", "--batch_size", "16", "--seed", "1"]
    stock = _run([sys.executable, SCRIPT, *args], {"NVIDIA_TF32_OVERRIDE": "0"})
    assert stock[0]["G"].startswith("torch.nn") and stock[0]["opt"].startswith("torch.optim")
    for prec, tol0, tol in (("fp32", 1e-4, 1e-2), ("bf16", 3e-2, 6e-2)):
        ours = _run([sys.executable, "-m", "eadgan_b200.run", SCRIPT, *args], {"EADGAN_PRECISION": prec})
        # the script really ran on the replacement modules / optimiser
        assert ours[0]["G"] == "eadgan_b200.nn" and ours[0]["opt"] == "eadgan_b200.optim", ours[0]
        for i, (a, b) in enumerate(zip(ours, stock)):
            t = tol0 if i == 0 else tol      # later iterations carry Adam's lr * sign(g) noise
            for k in ("g_loss", "d_loss"):
                assert abs(a[k] - b[k]) <= t * max(1.0, abs(b[k])), (prec, i, k, a, b)
