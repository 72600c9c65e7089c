"""Pre-training step of the MNIST affine approximator (MNIST/approximate_rpqmnxy.py:111-153): random codes -> affine
matrix R(theta) diag(p, q, 1) Skew(m, n) T(x, y) -> its top two rows [B, 6] -> 5-layer MLP -> MSE against the seven
affine parameters -> Adam(2e-4).  The reference composes the matrix on the HOST every iteration (:77-108, the same
eye(3) + slice-assignment pattern as the utils files); here it is composed on the device in closed form
(eadgan_b200.affine.mnist_matrix23), the Linear + LeakyReLU pairs run as fused kernels, MSE and Adam are the library's."""
from __future__ import annotations

import math

import torch

from .. import affine
from .. import nn as nn
from ..optim import Adam
from .mnist import AffineApproximator


def code_to_params(code):
    """from_latent_vector_2_affine_para, MNIST/approximate_rpqmnxy.py:57-72 (the operation order of the reference)"""
    return torch.stack((code[:, 0] * math.pi / 9, code[:, 1] * 0.2 + 1, code[:, 2] * 0.2 + 1, code[:, 3] * 0.2,
                        code[:, 4] * 0.2, code[:, 5] * 0.1, code[:, 6] * 0.1), dim=1)


class ApproximatorStep:
    def __init__(self, seed=0, device="cuda"):
        torch.manual_seed(seed)
        self.A = AffineApproximator().to(device)
        self.opt = Adam(self.A.parameters(), lr=0.0002, betas=(0.5, 0.999))     # :53
        self.mse = nn.MSELoss()

    def optimizers(self):
        return [self.opt]

    def __call__(self, code, record=None):
        """code [B, 7] in [-1, 1] on the device -> {"affine_loss"}"""
        rows = affine.mnist_matrix23(code).reshape(code.shape[0], 6)     # cat(A[:, 0], A[:, 1]) of :128
        self.opt.zero_grad()
        loss = self.mse(self.A(rows), code_to_params(code))
        loss.backward()
        if record is not None:
            record.append({"grads": [p.grad.detach().clone() for p in self.A.parameters()]})
        self.opt.step()
        if record is not None:
            record[-1]["params_after"] = [p.detach().clone() for p in self.A.parameters()]
        return {"affine_loss": loss.detach()}
