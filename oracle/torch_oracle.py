"""TEST INFRASTRUCTURE ONLY (the oracle).  Never imported by eadgan_b200/.

Stand-alone restatement, in stock ``torch.nn`` on CPU (or any torch device), of the
reference's adversarial training step.  It exists because /root/reference does not
travel to the GPU box; it is pinned, in the build container, against the reference
scripts themselves executed by oracle/ref_runner.py (tests/test_oracle_pin.py) and
against the committed fixtures in tests/golden/ (generated from the *reference*, see
oracle/make_golden.py).  The reference has no tests or golden vectors of its own
(SURVEY.md section 4), so those self-generated fixtures are the pin.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu-baseline / ``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import itertools
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

# --------------------------------------------------------------------------- #
# affine glue (utils_*.py)                                                    #
# --------------------------------------------------------------------------- #


def _eye_batch(n, like):
    return torch.eye(3, dtype=like.dtype, device=like.device).unsqueeze(0).repeat(n, 1, 1)


def _rot(theta):
    m = _eye_batch(theta.shape[0], theta)
    c, s = torch.cos(theta), torch.sin(theta)
    m[:, 0, 0] = c
    m[:, 0, 1] = -s
    m[:, 1, 0] = s
    m[:, 1, 1] = c
    return m


def _zoom(p, q):
    m = _eye_batch(p.shape[0], p)
    m[:, 0, 0] = p
    m[:, 1, 1] = q
    return m


def _shift(x, y):
    m = _eye_batch(x.shape[0], x)
    m[:, 0, 2] = x
    m[:, 1, 2] = y
    return m


def celeba_get_matrix(code5):
    """celebA/utils_rpqxy.py:25-38 (rescale) and :59-80 (R @ Z @ T)."""
    theta = code5[:, 0] * np.pi / 9
    p = code5[:, 1] * 0.2 + 1
    q = code5[:, 2] * 0.2 + 1
    x = code5[:, 3] * 0.1
    y = code5[:, 4] * 0.1
    return _rot(theta) @ _zoom(p, q) @ _shift(x, y)


def celeba_affine_regularizer(real_code, trans_code):
    """celebA/utils_rpqxy.py:82-116: closed-form recovery of the relative affine code."""
    rel = celeba_get_matrix(trans_code[:, :5]) @ torch.inverse(celeba_get_matrix(real_code[:, :5]))
    t1 = rel[:, 0, 0] * rel[:, 1, 0] - rel[:, 0, 1] * rel[:, 1, 1]
    t2 = rel[:, 0, 0] ** 2 + rel[:, 1, 1] ** 2 - rel[:, 0, 1] ** 2 - rel[:, 1, 0] ** 2
    th = 0.5 * torch.atan(2 * t1 / t2)
    p = rel[:, 0, 0] * torch.cos(th) + rel[:, 1, 0] * torch.sin(th)
    q = -rel[:, 0, 1] * torch.sin(th) + rel[:, 1, 1] * torch.cos(th)
    x = (rel[:, 0, 2] * torch.cos(th) + rel[:, 1, 2] * torch.sin(th)) / p
    y = (rel[:, 1, 2] * torch.cos(th) - rel[:, 0, 2] * torch.sin(th)) / q
    # celebA/utils_rpqxy.py:41-55 (inverse rescale)
    out = torch.stack((th / np.pi * 9, (p - 1) / 0.2, (q - 1) / 0.2, x / 0.1, y / 0.1), dim=1)
    return out.to(real_code.dtype)  # the reference's .float() (:116); dtype-preserving so an fp64 referee run works


def stn(img, theta23, padding_mode="border"):
    """transformation_2D.stn, celebA/EAD-GAN_celebA.py:149-153."""
    grid = F.affine_grid(theta23, img.size(), align_corners=False)
    return F.grid_sample(img, grid, padding_mode=padding_mode, align_corners=False)


# --------------------------------------------------------------------------- #
# CelebA (celebA/EAD-GAN_celebA.py)                                           #
# --------------------------------------------------------------------------- #


class CelebAGenerator(nn.Module):
    """celebA/EAD-GAN_celebA.py:67-102 (latent 200 + classes 10 + code 8 -> 3x64x64)."""

    def __init__(self, latent_dim=200, code_dim=8, n_classes=10, channels=3):
        super().__init__()
        d = latent_dim + code_dim + n_classes
        self.conv_blocks = nn.Sequential(
            nn.ConvTranspose2d(d, 1024, 4, 1, 0),
            nn.ConvTranspose2d(1024, 512, 4, stride=2, padding=1), nn.BatchNorm2d(512), nn.ReLU(),
            nn.ConvTranspose2d(512, 256, 4, stride=2, padding=1), nn.BatchNorm2d(256), nn.ReLU(),
            nn.ConvTranspose2d(256, 128, 4, stride=2, padding=1), nn.BatchNorm2d(128), nn.ReLU(),
            nn.ConvTranspose2d(128, channels, 4, stride=2, padding=1), nn.Tanh(),
        )

    def forward(self, noise, labels, code):
        x = torch.cat((noise, labels, code), -1)
        return self.conv_blocks(x.view(x.size(0), x.size(1), 1, 1))


class CelebADiscriminator(nn.Module):
    """celebA/EAD-GAN_celebA.py:105-138; channel 0 validity, 1..8 code, 9..18 class."""

    def __init__(self, code_dim=8, n_classes=10):
        super().__init__()
        self.code_dim, self.n_classes = code_dim, n_classes
        self.main = nn.Sequential(
            spectral_norm(nn.Conv2d(3, 128, 4, 2, 1)), nn.LeakyReLU(0.1, inplace=True),
            spectral_norm(nn.Conv2d(128, 256, 4, 2, 1)), nn.LeakyReLU(0.1, inplace=True),
            spectral_norm(nn.Conv2d(256, 512, 4, 2, 1)), nn.LeakyReLU(0.1, inplace=True),
            spectral_norm(nn.Conv2d(512, 1024, 4, 2, 1)), nn.LeakyReLU(0.1, inplace=True),
            nn.Conv2d(1024, 1 + n_classes + code_dim, 4, 1, 0),
        )

    def forward(self, img):
        out = self.main(img).squeeze()
        validity = torch.sigmoid(out[:, 0])
        cat = F.softmax(out[:, self.code_dim + 1: self.code_dim + 1 + self.n_classes], dim=1)
        cont = out[:, 1: self.code_dim + 1]
        return cat, cont, validity


def one_hot(labels, n, like):
    out = torch.zeros(labels.shape[0], n, dtype=like.dtype, device=like.device)
    out[torch.arange(labels.shape[0]), torch.as_tensor(labels, dtype=torch.long)] = 1.0
    return out


def sample_celeba(rs: np.random.RandomState, batch, latent_dim=200, code_dim=8, n_classes=10):
    """Host sampling in the reference's order, celebA/EAD-GAN_celebA.py:308-317."""
    z = rs.normal(0, 1, (batch, latent_dim))
    code = rs.uniform(-1, 1, (batch, code_dim))
    labels = rs.randint(0, n_classes, batch)
    return {"z": torch.tensor(z, dtype=torch.float32), "code": torch.tensor(code, dtype=torch.float32),
            "labels": torch.tensor(labels, dtype=torch.long)}


def build_celeba(seed=0, device="cpu", dtype=torch.float32):
    """Models + the three Adams of celebA/EAD-GAN_celebA.py:172-173,211-217.
    Construction order (G then D) after ``torch.manual_seed`` fixes the random init."""
    torch.manual_seed(seed)
    G, D = CelebAGenerator(), CelebADiscriminator()
    G.to(device=device, dtype=dtype)
    D.to(device=device, dtype=dtype)
    betas = (0.5, 0.999)
    st = {
        "G": G, "D": D,
        "opt_G": torch.optim.Adam(G.parameters(), lr=0.001, betas=betas),
        "opt_D": torch.optim.Adam(D.parameters(), lr=0.0002, betas=betas),
        "opt_info": torch.optim.Adam(itertools.chain(G.parameters(), D.parameters()), lr=0.0002, betas=betas),
    }
    return st


def _snap(opt):
    ps = [p for g in opt.param_groups for p in g["params"]]
    return [None if p.grad is None else p.grad.detach().clone() for p in ps]


def _params(opt):
    return [p.detach().clone() for g in opt.param_groups for p in g["params"]]


def _state(G, D):
    """full module state (weights, BN running stats, spectral-norm u/v) after a phase: lets a test restart
    the next phase of another implementation from THIS run's state (SURVEY.md section 7.3-1 iv)."""
    return {"G": {k: v.detach().clone() for k, v in G.state_dict().items()},
            "D": {k: v.detach().clone() for k, v in D.state_dict().items()}}


def step_celeba(st, imgs, draws, record=True):
    """One iteration of celebA/EAD-GAN_celebA.py:297-401 on batch ``imgs`` [B,3,64,64]."""
    G, D = st["G"], st["D"]
    dev, dt = imgs.device, imgs.dtype
    B = imgs.shape[0]
    bce, mse, ce = nn.BCELoss(), nn.MSELoss(), nn.CrossEntropyLoss()
    valid = torch.ones(B, device=dev, dtype=dt)            # :302
    fake = torch.zeros(B, device=dev, dtype=dt)            # :303
    z = draws["z"].to(dev, dt)
    code = draws["code"].to(dev, dt)
    labels = draws["labels"].to(dev)
    label_input = one_hot(labels, 10, z)
    imgs = imgs.to(dt)
    A = celeba_get_matrix(code[:, :5])                      # :325
    scaled = stn(imgs, A[:, 0:2])                           # :327
    rec = {"phases": []}

    # ---- phase G  (:334-345)
    st["opt_G"].zero_grad()
    gen = G(z, label_input, code)
    _, _, validity = D(gen)
    g_loss = bce(validity, valid)
    g_loss.backward()
    if record:
        rec["phases"].append({"name": "G", "grads": _snap(st["opt_G"])})
    st["opt_G"].step()
    if record:
        rec["phases"][-1]["params_after"] = _params(st["opt_G"])
        rec["phases"][-1]["state_after"] = _state(G, D)

    # ---- phase D  (:353-366)
    st["opt_D"].zero_grad()
    _, _, real_pred = D(scaled)
    d_real = bce(real_pred, valid)
    _, _, fake_pred = D(gen.detach())
    d_fake = bce(fake_pred, fake)
    d_loss = (d_real + d_fake) / 2
    d_loss.backward()
    if record:
        rec["phases"].append({"name": "D", "grads": _snap(st["opt_D"])})
    st["opt_D"].step()
    if record:
        rec["phases"][-1]["params_after"] = _params(st["opt_D"])
        rec["phases"][-1]["state_after"] = _state(G, D)

    # ---- phase info  (:375-401)
    st["opt_info"].zero_grad()
    gen = G(z, label_input, code)
    pred_label, pred_code, _ = D(gen)
    info1 = ce(pred_label, labels) + mse(pred_code, code)   # CE on softmax output (:383)
    _, transform_code, _ = D(scaled)
    _, real_code, _ = D(imgs)
    pred_aff = celeba_affine_regularizer(real_code, transform_code)
    info_loss = info1 + mse(pred_aff, code[:, :5])
    info_loss.backward()
    if record:
        rec["phases"].append({"name": "info", "grads": _snap(st["opt_info"])})
    st["opt_info"].step()
    if record:
        rec["phases"][-1]["params_after"] = _params(st["opt_info"])

    rec["losses"] = {"g_loss": g_loss.item(), "d_loss": d_loss.item(), "info_loss": info_loss.item()}
    return rec


# --------------------------------------------------------------------------- #
# synthetic inputs (SURVEY.md section 8 d)                                    #
# --------------------------------------------------------------------------- #


def synth_celeba_images(batch, seed=0):
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand(batch, 3, 64, 64, generator=g) * 2 - 1


def summarize(t: torch.Tensor):
    """Size-independent fingerprint of a tensor used by the golden fixtures."""
    t = t.detach().double().flatten()
    idx = torch.linspace(0, t.numel() - 1, steps=min(8, t.numel())).long()
    return {"sum": float(t.sum()), "l2": float(t.norm()), "absmax": float(t.abs().max()),
            "probe": [float(v) for v in t[idx]]}
