"""dSprites stage-2 training step (dSprites/rp.py): frozen Encoder_pxy alignment, Discriminator, Generator,
Encoder; two optimisation phases (D / info) and two Adams (rp.py:276-282; ``optimizer_G`` is created by the
reference but never stepped).

Modules are written as the reference writes them (rp.py:56-190) against ``eadgan_b200.nn``, so the
state_dict layout is the reference's and every conv stack runs on the sm_100a kernels (the 32/64-channel
k4 s2 p1 trunks through the tcgen05 chain in bf16 mode, the Linear heads through the fp32 SIMT GEMM).

Deviation (benign): the frozen Encoder_pxy is run without building an autograd
graph.  The reference leaves it grad-tracked, which only computes gradients nobody reads (its parameters
belong to no optimiser); every gradient that IS consumed is unchanged.
"""
from __future__ import annotations

import itertools

import torch

from .. import affine, chain, functional as Fn
from .. import nn as nn
from ..optim import Adam
from .._lib import ACT_SIGMOID

N_CLASSES, CODE = 3, 4


def _trunk(cin, slope, sn):
    wrap = nn.spectral_norm if sn else (lambda m: m)
    layers = []
    for a, b in ((cin, 32), (32, 32), (32, 64), (64, 64)):
        layers += [wrap(nn.Conv2d(a, b, 4, 2, 1)), nn.LeakyReLU(slope, inplace=True)]
    return nn.Sequential(*layers)


class Encoder_pxy(torch.nn.Module):
    def __init__(self, channels=1, out_dim=3):
        super().__init__()
        self.conv_block = _trunk(channels, 0.1, False)
        self.fc1 = nn.Linear(1024, out_dim)

    def forward(self, img):
        x = self.conv_block(img)
        return self.fc1(x.reshape(x.shape[0], -1))


class Discriminator(torch.nn.Module):
    def __init__(self, channels=1):
        super().__init__()
        self.conv_block = _trunk(channels, 0.2, True)
        self.fc1 = nn.Sequential(nn.spectral_norm(nn.Linear(1024, 128)), nn.LeakyReLU(0.2, inplace=True))
        self.fc2 = nn.Linear(128, 1)

    def forward(self, img):
        x = self.conv_block(img)
        x = self.fc1(x.reshape(x.shape[0], -1))
        return Fn.activation(self.fc2(x), ACT_SIGMOID)


class Generator(torch.nn.Module):
    def __init__(self, channels=1, code_dim=CODE):
        super().__init__()
        blk = []
        for _ in range(3):
            blk += [nn.ConvTranspose2d(64, 64, 4, 2, 1), nn.BatchNorm2d(64), nn.ReLU()]
        blk.append(nn.ConvTranspose2d(64, channels, 4, 2, 1))
        self.conv_block = nn.Sequential(*blk)                 # registration order of rp.py:128-146
        self.fc1 = nn.Sequential(nn.Linear(N_CLASSES + code_dim, 128), nn.ReLU())
        self.fc2 = nn.Sequential(nn.Linear(128, 64 * 4 * 4), nn.ReLU())

    def forward(self, z_c):
        x = self.fc2(self.fc1(z_c))
        x = self.conv_block(x.view(x.shape[0], 64, 4, 4))
        return Fn.activation(x, ACT_SIGMOID)


class Encoder(torch.nn.Module):
    def __init__(self, channels=1, code_dim=CODE):
        super().__init__()
        self.conv_block = _trunk(channels, 0.2, True)
        self.fc1 = nn.Sequential(nn.spectral_norm(nn.Linear(1024, 128)), nn.LeakyReLU(0.2, inplace=True))
        self.fc2 = nn.Sequential(nn.spectral_norm(nn.Linear(128, 128)), nn.LeakyReLU(0.2, inplace=True))
        self.cat_layer = nn.Sequential(nn.spectral_norm(nn.Linear(128, N_CLASSES)), nn.Softmax())
        self.cont_layer = nn.Sequential(nn.spectral_norm(nn.Linear(128, code_dim)))

    def forward(self, img):
        x = self.conv_block(img)
        x = self.fc2(self.fc1(x.reshape(x.shape[0], -1)))
        return self.cat_layer(x), self.cont_layer(x)


class DSpritesStep:
    """Owns Encoder_pxy (frozen), E, D, G, the two Adams and the losses; ``__call__`` runs one iteration."""

    def __init__(self, seed=0, device="cuda", pxy_state=None):
        torch.manual_seed(seed)  # construction order of rp.py:255-258: encoder_pxy, encoder, discriminator, generator
        self.Epxy, self.E, self.D, self.G = Encoder_pxy(), Encoder(), Discriminator(), Generator()
        if pxy_state is not None:
            self.Epxy.load_state_dict(pxy_state)              # rp.py:271-273 (encoder_pxy_50000.pt)
        self.Epxy.eval()
        for m in (self.Epxy, self.E, self.D, self.G):
            m.to(device)
        betas = (0.5, 0.999)
        self.opt_D = Adam(self.D.parameters(), lr=0.0002, betas=betas)                                  # :277
        self.opt_info = Adam(itertools.chain(self.G.parameters(), self.E.parameters()), lr=0.0001, betas=betas)
        self.bce, self.mse = nn.BCELoss(), nn.MSELoss()
        self.device = torch.device(device)

    def optimizers(self):
        return [self.opt_D, self.opt_info]

    def _aligned(self, img):
        with torch.no_grad():
            code = self.Epxy(img)
            return affine.stn(img, affine.dsprites_align_inverse(code))

    @staticmethod
    def _snap(opt, rec, name):
        if rec is not None:
            ps = [p for g in opt.param_groups for p in g["params"]]
            rec.append({"name": name, "grads": [None if p.grad is None else p.grad.detach().clone() for p in ps]})

    @staticmethod
    def _after(opt, rec):
        if rec is not None:
            rec[-1]["params_after"] = [p.detach().clone() for g in opt.param_groups for p in g["params"]]

    def __call__(self, img_u8, code_d, labels_d, code_info, labels_info, record=None, after_phase=None):
        """img_u8 uint8 [B,64,64]; code_* [B,4] in [-1,1]; labels_* [B] int64 -- all on the device."""
        E, D, G = self.E, self.D, self.G
        B = img_u8.shape[0]
        img = img_u8.unsqueeze(1).float()
        valid = torch.ones(B, 1, device=img.device)
        fake = torch.zeros(B, 1, device=img.device)

        def onehot(labels):
            o = torch.zeros(B, N_CLASSES, device=img.device)
            o.scatter_(1, labels.view(-1, 1), 1.0)
            return o

        # spectral-norm power iterations of the coming forwards, issued early on a side stream (they depend only on
        # weights that stay fixed until the owning optimiser steps): D twice in phase D, E three times in phase info
        chain.clear_prefetch(D.conv_block)          # leftovers of an aborted step, if any
        chain.clear_prefetch(E.conv_block)
        chain.prefetch_spectral_norm(D.conv_block, 2)
        chain.prefetch_spectral_norm(E.conv_block, 3)

        # phase D -- rp.py:379-419
        align_img = self._aligned(img)
        trans_img = affine.stn(align_img, affine.dsprites_matrix23(code_d))
        gen = G(torch.cat((onehot(labels_d), code_d), dim=1))
        d_real = D(trans_img)                      # real first, then fake: the spectral-norm u, v advance per call
        d_fake = D(gen.detach())
        d_loss = (self.bce(d_fake, fake) + self.bce(d_real, valid)) / 2
        self.opt_D.zero_grad()
        d_loss.backward()
        self._snap(self.opt_D, record, "D")
        self.opt_D.step()
        self._after(self.opt_D, record)
        if after_phase is not None:
            after_phase(0)
        # the info phase differentiates THROUGH D (g_loss) but opt_info owns only G and E: D is frozen for it, its
        # weight gradients (computed and never read by the reference) are not launched
        chain.set_trainable(D, False)
        chain.prefetch_spectral_norm(D.conv_block, 1)      # D(gen) of the info phase

        # phase info -- rp.py:424-482
        lab = onehot(labels_info)
        gen = G(torch.cat((lab, code_info), dim=1))
        rec_cat, rec_cont = E(gen)
        g_loss = self.bce(D(gen), valid)
        cat_loss = Fn.mutual_info_loss(rec_cat, lab)
        cont_loss = self.mse(rec_cont, code_info)
        align_img = self._aligned(img)
        trans_img = affine.stn(align_img, affine.dsprites_matrix23(code_info))
        align_cat, align_cont = E(align_img)
        trans_cat, trans_cont = E(trans_img)
        affine_loss = self.mse(affine.dsprites_relative_code(align_cont, trans_cont), code_info)
        rel_cat_loss = Fn.mutual_info_loss(trans_cat, align_cat.detach())
        total = cat_loss + cont_loss + affine_loss + g_loss + rel_cat_loss
        self.opt_info.zero_grad()
        total.backward()
        chain.set_trainable(D, True)
        self._snap(self.opt_info, record, "info")
        self.opt_info.step()
        self._after(self.opt_info, record)
        return {"d_loss": d_loss.detach(), "g_loss": g_loss.detach(), "cat_loss": cat_loss.detach(),
                "cont_loss": cont_loss.detach(), "affine_loss": affine_loss.detach(),
                "relative_cat_loss": rel_cat_loss.detach(), "total": total.detach()}
