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


def _run(cmd, env_extra, cwd=ROOT):
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), **env_extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400, cwd=cwd, env=env)
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


def _compare(ours, stock, keys, tol0, tol, what):
    for i, (a, b) in enumerate(zip(ours, stock)):
        t = tol0 if i == 0 else tol          # later iterations carry Adam's lr * sign(g) noise
        for k in keys:
            assert abs(a[k] - b[k]) <= t * max(1.0, abs(b[k])), (what, i, k, a, b)


def test_rp_shaped_script_with_shadowed_utils_and_grid_sample_backward(cuda):
    """dSprites/rp.py-shaped script (tests/scripts/dSprites/rp_mini.py): a grad-tracked frozen alignment encoder ->
    torch.inverse -> F.affine_grid / F.grid_sample BACKWARD -> the conv chain's input gradient, helper functions
    star-imported from utils_rp / utils_pxy.  Stock run: the oracle's host-style helpers + ATen; shim run: the
    shadow modules (eadgan_b200/shadow/dSprites), our affine_grid / grid_sample kernels, closed-form 3x3 inverse."""
    script = os.path.join(ROOT, "tests", "scripts", "dSprites", "rp_mini.py")
    args = ["--n_iter", "3", "--batch_size", "16", "--seed", "2"]
    stock = _run([sys.executable, script, *args], {"NVIDIA_TF32_OVERRIDE": "0"})
    assert stock[0]["G"].startswith("torch.nn") and stock[0]["utils"] == "oracle.torch_oracle"
    assert stock[0]["pxy_grad_l1"] > 0            # the frozen encoder DOES receive gradients through grid_sample
    keys = ("d_loss", "g_loss", "info_loss", "affine_loss", "total")
    for prec, tol0, tol in (("fp32", 1e-4, 1e-2), ("bf16", 3e-2, 6e-2)):
        ours = _run([sys.executable, "-m", "eadgan_b200.run", script, *args], {"EADGAN_PRECISION": prec})
        assert ours[0]["G"] == "eadgan_b200.nn" and ours[0]["opt"] == "eadgan_b200.optim", ours[0]
        assert ours[0]["utils"] == "utils_rp", ours[0]          # the shadow module, not the file next to the script
        _compare(ours, stock, keys, tol0, tol, prec)
        rel = abs(ours[0]["pxy_grad_l1"] - stock[0]["pxy_grad_l1"]) / stock[0]["pxy_grad_l1"]
        assert rel <= (2e-3 if prec == "fp32" else 0.1), (prec, ours[0]["pxy_grad_l1"], stock[0]["pxy_grad_l1"])


def test_mnist_shaped_script(cuda, tmp_path):
    """MNIST/EAD-GAN_rpqmnxy.py-shaped script (tests/scripts/MNIST/mnist_mini.py): Linear -> view -> BatchNorm2d,
    Upsample, BatchNorm2d(C, 0.8), nn.Softmax(), LSGAN loss, the approximator-based affine regulariser loaded from
    rpqmnxy_approximator.pt in the working directory (a seeded random-init stand-in, as in the oracle)."""
    import torch
    from oracle import torch_oracle as O
    torch.save(O.mnist_approximator_state(3), os.path.join(tmp_path, "rpqmnxy_approximator.pt"))
    script = os.path.join(ROOT, "tests", "scripts", "MNIST", "mnist_mini.py")
    args = ["--n_iter", "3", "--batch_size", "16", "--seed", "3"]
    stock = _run([sys.executable, script, *args], {"NVIDIA_TF32_OVERRIDE": "0"}, cwd=str(tmp_path))
    assert stock[0]["G"].startswith("torch.nn") and stock[0]["utils"] == "oracle.torch_oracle"
    for prec, tol0, tol in (("fp32", 1e-4, 1e-2), ("bf16", 3e-2, 6e-2)):
        ours = _run([sys.executable, "-m", "eadgan_b200.run", script, *args], {"EADGAN_PRECISION": prec}, cwd=str(tmp_path))
        assert ours[0]["G"] == "eadgan_b200.nn" and ours[0]["utils"] == "utils_rpqmnxy", ours[0]
        _compare(ours, stock, ("g_loss", "d_loss", "info_loss"), tol0, tol, prec)
