"""MNIST training step (MNIST/EAD-GAN_rpqmnxy.py:337-446; BASELINE configs[0]): Generator (Linear + Upsample +
3x3 convs + BatchNorm eps 0.8), spectral-normalised Discriminator and Encoder (3x3 stride-2 convs), LSGAN (MSE)
adversarial loss, three optimisation phases (G / D / info) and three Adams; the relative-affine code is
recovered by the frozen 5-layer MLP approximator of MNIST/utils_rpqmnxy.py:12-43.

Module classes are written as the reference writes them (:71-175) against ``eadgan_b200.nn``: same
constructor arguments, registration order and state_dict layout; ``weights_init_normal`` (:54-60) works on
them unchanged because the class names keep the ``Conv`` / ``BatchNorm`` substrings.  None of these layers is
k4 s2 p1, so every conv runs on the any-geometry fp32 SIMT kernels (csrc/simt_conv.cu) in both precision
modes; BASELINE lists this configuration as the reference's own CPU-runnable case.

Deviation (benign): the approximator's parameters are frozen with
``requires_grad_(False)``.  The reference sets a meaningless attribute instead (utils_rpqmnxy.py:43), so it
computes parameter gradients nobody reads; the gradient THROUGH the approximator to the encoder is unchanged.
"""
from __future__ import annotations

import itertools

import torch

from .. import affine
from .. import nn as nn
from ..optim import Adam

LATENT, CODE, CLASSES, IMG, CHANNELS = 62, 7, 10, 32, 1   # argparse defaults, MNIST/EAD-GAN_rpqmnxy.py:42-46


def weights_init_normal(m):
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find("BatchNorm") != -1:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


class AffineApproximator(torch.nn.Module):
    def __init__(self):
        super().__init__()
        layers = [nn.Linear(6, 256), nn.LeakyReLU()]
        for _ in range(3):
            layers += [nn.Linear(256, 256), nn.LeakyReLU()]
        layers.append(nn.Linear(256, 7))
        self.fc_block = nn.Sequential(*layers)

    def forward(self, x):
        return self.fc_block(x)


class Generator(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.init_size = IMG // 4
        self.l1 = nn.Sequential(nn.Linear(LATENT + CLASSES + CODE, 128 * self.init_size ** 2))
        self.conv_blocks = nn.Sequential(
            nn.BatchNorm2d(128), nn.Upsample(scale_factor=2), nn.Conv2d(128, 128, 3, stride=1, padding=1),
            nn.BatchNorm2d(128, 0.8), nn.LeakyReLU(0.2, inplace=True), nn.Upsample(scale_factor=2),
            nn.Conv2d(128, 64, 3, stride=1, padding=1), nn.BatchNorm2d(64, 0.8), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(64, CHANNELS, 3, stride=1, padding=1), nn.Tanh())

    def forward(self, noise, labels, code):
        out = self.l1(torch.cat((noise, labels, code), -1))
        return self.conv_blocks(out.view(out.shape[0], 128, self.init_size, self.init_size))


def _trunk(bn):
    layers = []
    for i, (a, b) in enumerate(((CHANNELS, 16), (16, 32), (32, 64), (64, 128))):
        layers += [nn.spectral_norm(nn.Conv2d(a, b, 3, 2, 1)), nn.LeakyReLU(0.2, inplace=True)]
        if bn and i > 0:
            layers.append(nn.BatchNorm2d(b, 0.8))
    return nn.Sequential(*layers)


class Discriminator(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv_blocks = _trunk(False)
        self.adv_layer = nn.Sequential(nn.spectral_norm(nn.Linear(128 * (IMG // 16) ** 2, 1)))

    def forward(self, img):
        out = self.conv_blocks(img)
        return self.adv_layer(out.reshape(out.shape[0], -1))


class Encoder(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.conv_blocks = _trunk(True)
        feat = 128 * (IMG // 16) ** 2
        self.aux_layer = nn.Sequential(nn.spectral_norm(nn.Linear(feat, CLASSES)), nn.Softmax())
        self.latent_layer = nn.Sequential(nn.spectral_norm(nn.Linear(feat, CODE)))
        self.noise_layer = nn.Sequential(nn.spectral_norm(nn.Linear(feat, LATENT)))

    def forward(self, img):
        out = self.conv_blocks(img)
        out = out.reshape(out.shape[0], -1)
        return self.aux_layer(out), self.latent_layer(out), self.noise_layer(out)


class MnistStep:
    """Owns the approximator (frozen), G, D, E, the three Adams and the losses; ``__call__`` runs one iteration."""

    def __init__(self, seed=0, device="cuda", approximator_state=None):
        torch.manual_seed(seed)   # construction order: approximator (utils import), G, D, E, then the init pass
        self.A = AffineApproximator()
        if approximator_state is not None:
            self.A.load_state_dict(approximator_state)       # utils_rpqmnxy.py:36-41 (rpqmnxy_approximator.pt)
        self.A.eval()
        self.A.requires_grad_(False)
        self.G, self.D, self.E = Generator(), Discriminator(), Encoder()
        for m in (self.G, self.D, self.E):
            m.apply(weights_init_normal)                     # :229-231
        for m in (self.A, self.G, self.D, self.E):
            m.to(device)
        betas, lr = (0.5, 0.999), 0.0001
        self.opt_G = Adam([{"params": self.G.parameters()}], lr=lr, betas=betas)              # :249
        self.opt_D = Adam(self.D.parameters(), lr=lr * 2, betas=betas)                        # :250
        self.opt_info = Adam(itertools.chain(self.G.parameters(), self.E.parameters()), lr=lr, betas=betas)
        self.mse, self.ce = nn.MSELoss(), nn.CrossEntropyLoss()
        self.device = torch.device(device)

    def optimizers(self):
        return [self.opt_G, self.opt_D, self.opt_info]

    @staticmethod
    def _snap(opt, rec, name):
        if rec is not None:
            ps = [p for g in opt.param_groups for p in g["params"]]
            rec.append({"name": name, "grads": [None if p.grad is None else p.grad.detach().clone() for p in ps]})

    @staticmethod
    def _after(opt, rec):
        if rec is not None:
            rec[-1]["params_after"] = [p.detach().clone() for g in opt.param_groups for p in g["params"]]

    def __call__(self, imgs, z, code, labels, record=None, after_phase=None):
        """imgs [B,1,32,32] in [-1,1]; z [B,62]; code [B,7]; labels [B] int64 -- all on the device."""
        G, D, E = self.G, self.D, self.E
        B = imgs.shape[0]
        valid = torch.ones(B, 1, device=imgs.device)
        fake = torch.zeros(B, 1, device=imgs.device)
        onehot = torch.zeros(B, CLASSES, device=imgs.device)
        onehot.scatter_(1, labels.view(-1, 1), 1.0)
        scaled = affine.stn(imgs, affine.mnist_matrix23(code))                 # :363-365

        # phase G -- :373-386
        self.opt_G.zero_grad()
        gen = G(z, onehot, code)
        g_loss = self.mse(D(gen), valid)
        g_loss.backward()
        self._snap(self.opt_G, record, "G")
        self.opt_G.step()
        self._after(self.opt_G, record)
        if after_phase is not None:
            after_phase(0)

        # phase D -- :393-407
        self.opt_D.zero_grad()
        d_loss = (self.mse(D(scaled), valid) + self.mse(D(gen.detach()), fake)) / 2
        d_loss.backward()
        self._snap(self.opt_D, record, "D")
        self.opt_D.step()
        self._after(self.opt_D, record)
        if after_phase is not None:
            after_phase(1)

        # phase info -- :413-446  (lambda_cat 1, lambda_con 0.1, lambda_affine 0.1, :201-203)
        self.opt_info.zero_grad()
        gen = G(z, onehot, code)
        pred_label, pred_code, _ = E(gen)
        info = 1 * self.ce(pred_label, labels) + 0.1 * self.mse(pred_code, code)
        _, transform_code, _ = E(scaled)
        _, real_code, _ = E(imgs)
        pred = affine.mnist_code_from_params(self.A(affine.mnist_relative_rows(real_code, transform_code)))
        info = info + 0.1 * self.mse(pred, code)
        info.backward()
        self._snap(self.opt_info, record, "info")
        self.opt_info.step()
        self._after(self.opt_info, record)
        return {"g_loss": g_loss.detach(), "d_loss": d_loss.detach(), "info_loss": info.detach()}
