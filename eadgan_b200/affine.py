"""Affine-code glue of the training step: latent code -> 3x3 matrix -> STN warp, and the
closed-form recovery of relative affine parameters (reference: celebA/utils_rpqxy.py:25-116,
dSprites/utils_rp.py:23-147, dSprites/utils_pxy.py:24-126).

SURVEY.md section 8(f) ranks these as "next" rows: they stay stock differentiable torch
ops for now, but are restated here to run ENTIRELY ON THE DEVICE -- the reference builds
``torch.eye(3)`` on the CPU and assigns CUDA slices into it, i.e. 8-9 synchronising D2H
copies per call; this version has no host round trip.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as TF


def _mat3(rows):
    """rows: 3x3 nested list of [B] tensors / python floats -> [B,3,3]."""
    ref = next(e for r in rows for e in r if torch.is_tensor(e))
    out = []
    for r in rows:
        out.append(torch.stack([e if torch.is_tensor(e) else torch.full_like(ref, float(e)) for e in r], dim=1))
    return torch.stack(out, dim=1)


def compose_rzt(theta, p, q, x, y):
    """R(theta) @ diag(p,q,1) @ T(x,y) written out (celebA/utils_rpqxy.py:64-77)."""
    c, s = torch.cos(theta), torch.sin(theta)
    return _mat3([[c * p, -s * q, c * p * x - s * q * y],
                  [s * p, c * q, s * p * x + c * q * y],
                  [0.0, 0.0, 1.0]])


def celeba_matrix(code5):
    return compose_rzt(code5[:, 0] * (math.pi / 9), code5[:, 1] * 0.2 + 1, code5[:, 2] * 0.2 + 1,
                       code5[:, 3] * 0.1, code5[:, 4] * 0.1)


def celeba_relative_code(real_code, trans_code):
    """affine_regularzier of celebA/utils_rpqxy.py:82-116."""
    rel = celeba_matrix(trans_code[:, :5]) @ torch.linalg.inv(celeba_matrix(real_code[:, :5]))
    a, b, c, d = rel[:, 0, 0], rel[:, 0, 1], rel[:, 1, 0], rel[:, 1, 1]
    th = 0.5 * torch.atan(2 * (a * c - b * d) / (a * a + d * d - b * b - c * c))
    ct, st = torch.cos(th), torch.sin(th)
    p = a * ct + c * st
    q = -b * st + d * ct
    x = (rel[:, 0, 2] * ct + rel[:, 1, 2] * st) / p
    y = (rel[:, 1, 2] * ct - rel[:, 0, 2] * st) / q
    return torch.stack((th * (9 / math.pi), (p - 1) / 0.2, (q - 1) / 0.2, x / 0.1, y / 0.1), dim=1)


def stn(img, theta23, padding_mode="border"):
    """transformation_2D.stn (celebA/EAD-GAN_celebA.py:149-153); stock torch op ("next" row f2)."""
    grid = TF.affine_grid(theta23, list(img.shape), align_corners=False)
    return TF.grid_sample(img, grid, padding_mode=padding_mode, align_corners=False)
