"""colored-dSprites stage-2 training step (colored_dSprites/rp_color.py:362-516; BASELINE configs[2], the
data-parallel configuration): the dSprites step with 3-channel images, a 7-d code (4 affine entries + 3
colour gains), an Encoder_pxy that also emits 3 colour-alignment gains, and both Adams at lr 2e-4
(rp_color.py:39,275-280).

Module classes are the dSprites ones (the two scripts differ only in channel / code widths:
colored_dSprites/rp_color.py:59-192 vs dSprites/rp.py:61-194) built against ``eadgan_b200.nn``.  The colour
arithmetic (utils_rp_color.py:38-47, utils_pxy.py:48-57) is elementwise glue on ``[B,3]`` gains.

Deviation (benign, as in steps/dsprites.py): the frozen Encoder_pxy runs without an autograd graph.
"""
from __future__ import annotations

import itertools

import torch

from .. import affine, chain, functional as Fn
from .. import nn as nn
from ..optim import Adam
from .dsprites import Discriminator, DSpritesStep, Encoder, Encoder_pxy, Generator, N_CLASSES

CODE, CHANNELS = 7, 3


def colorize(img_u8, gains):
    """uint8 [B,64,64] x float64 gains [B,3,1,1] -> float32 [B,3,64,64]  (rp_color.py:366-381: the product
    is formed in float64 and then cast, which is what fixes the rounding of every pixel)."""
    return (img_u8.unsqueeze(1).repeat(1, 3, 1, 1) * gains.double()).float()


class ColoredDSpritesStep(DSpritesStep):
    """Owns Encoder_pxy (frozen), E, D, G, the two Adams and the losses; ``__call__`` runs one iteration."""

    def __init__(self, seed=0, device="cuda", pxy_state=None):
        torch.manual_seed(seed)  # construction order of rp_color.py:253-256
        self.Epxy, self.E = Encoder_pxy(CHANNELS, 6), Encoder(CHANNELS, CODE)
        self.D, self.G = Discriminator(CHANNELS), Generator(CHANNELS, CODE)
        if pxy_state is not None:
            self.Epxy.load_state_dict(pxy_state)              # rp_color.py:269-271 (encoder_pxy_color_50000.pt)
        self.Epxy.eval()
        for m in (self.Epxy, self.E, self.D, self.G):
            m.to(device)
        betas = (0.5, 0.999)
        self.opt_D = Adam(self.D.parameters(), lr=0.0002, betas=betas)
        self.opt_info = Adam(itertools.chain(self.G.parameters(), self.E.parameters()), lr=0.0002, betas=betas)
        self.bce, self.mse = nn.BCELoss(), nn.MSELoss()
        self.device = torch.device(device)

    def _aligned(self, img):
        with torch.no_grad():
            code = self.Epxy(img)
            gains = (code[:, 3:] * 0.1 + 1).unsqueeze(2).unsqueeze(3)
            return affine.stn(img, affine.dsprites_align_inverse(code)) / gains           # rp_color.py:385-394

    @staticmethod
    def _distorted(align_img, code):
        gains = (code[:, 4:] * 0.5 + 1).unsqueeze(2).unsqueeze(3)
        return affine.stn(align_img, affine.dsprites_matrix23(code)) * gains              # rp_color.py:416-424

    def __call__(self, img_u8, gains, code_d, labels_d, code_info, labels_info, record=None, after_phase=None):
        """img_u8 uint8 [B,64,64]; gains float64 [B,3,1,1] in [0.5,1]; code_* [B,7] in [-1,1]; labels_* [B]
        int64 -- all on the device."""
        E, D, G = self.E, self.D, self.G
        B = img_u8.shape[0]
        img = colorize(img_u8, gains)
        valid = torch.ones(B, 1, device=img.device)
        fake = torch.zeros(B, 1, device=img.device)

        def onehot(labels):
            o = torch.zeros(B, N_CLASSES, device=img.device)
            o.scatter_(1, labels.view(-1, 1), 1.0)
            return o

        chain.clear_prefetch(D.conv_block)          # leftovers of an aborted step, if any
        chain.clear_prefetch(E.conv_block)
        chain.prefetch_spectral_norm(D.conv_block, 2)      # see steps/dsprites.py
        chain.prefetch_spectral_norm(E.conv_block, 3)

        # phase D -- rp_color.py:397-441
        align_img = self._aligned(img)
        trans_img = self._distorted(align_img, code_d)
        gen = G(torch.cat((onehot(labels_d), code_d), dim=1))
        d_real = D(trans_img)
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

        # phase info -- rp_color.py:444-516
        lab = onehot(labels_info)
        gen = G(torch.cat((lab, code_info), dim=1))
        rec_cat, rec_cont = E(gen)
        g_loss = self.bce(D(gen), valid)
        cat_loss = Fn.mutual_info_loss(rec_cat, lab)
        cont_loss = self.mse(rec_cont, code_info)
        align_img = self._aligned(img)
        trans_img = self._distorted(align_img, code_info)
        align_cat, align_cont = E(align_img)
        trans_cat, trans_cont = E(trans_img)
        affine_loss = self.mse(affine.colored_relative_code(align_cont, trans_cont), code_info)
        rel_cat_loss = Fn.mutual_info_loss(trans_cat, align_cat.detach())
        total = cat_loss + cont_loss + affine_loss + rel_cat_loss + g_loss
        self.opt_info.zero_grad()
        total.backward()
        chain.set_trainable(D, True)
        self._snap(self.opt_info, record, "info")
        self.opt_info.step()
        self._after(self.opt_info, record)
        return {"d_loss": d_loss.detach(), "g_loss": g_loss.detach(), "cat_loss": cat_loss.detach(),
                "cont_loss": cont_loss.detach(), "affine_loss": affine_loss.detach(),
                "relative_cat_loss": rel_cat_loss.detach(), "total": total.detach()}
