"""Stage-1 pre-training step of the alignment encoder (dSprites/pxy.py:156-187; colored variant
colored_dSprites/pxy_color.py:162-216): Encoder_pxy forward on an image and on its affinely (and, colored,
chromatically) distorted copy, closed-form relative-code recovery, one MSE loss, one Adam.  SURVEY.md
section 8(f) rank 4: same operator set as the stage-2 steps; the conv trunk runs on the tcgen05 chain in bf16
mode, and the STN backward is NOT needed here (the image is a leaf without gradient)."""
from __future__ import annotations

import torch

from .. import affine
from .. import nn as nn
from ..optim import Adam
from .colored import colorize
from .dsprites import Encoder_pxy


class PxyStep:
    def __init__(self, seed=0, device="cuda", colored=False):
        torch.manual_seed(seed)
        self.colored = colored
        self.E = Encoder_pxy(3, 6) if colored else Encoder_pxy()
        self.E.to(device)
        self.opt_E = Adam(self.E.parameters(), lr=0.0002, betas=(0.5, 0.999))     # pxy.py:37,122
        self.mse = nn.MSELoss()
        self.device = torch.device(device)

    def optimizers(self):
        return [self.opt_E]

    def __call__(self, img_u8, code, gains=None, record=None):
        """img_u8 uint8 [B,64,64]; code [B,3] (colored: [B,6]) in [-1,1]; gains float64 [B,3,1,1] (colored)."""
        if self.colored:
            img = colorize(img_u8, gains)
        else:
            img = img_u8.unsqueeze(1).float()
        real_code = self.E(img)
        # pxy_color.py:90 warps with zero padding, every other script with border padding
        trans = affine.stn(img, affine.pxy_matrix23(code), padding_mode="zeros" if self.colored else "border")
        if self.colored:
            trans = trans * (code[:, 3:] * 0.1 + 1).unsqueeze(2).unsqueeze(3)
        trans_code = self.E(trans)
        loss = self.mse(affine.pxy_relative_code(real_code, trans_code), code)
        self.opt_E.zero_grad()
        loss.backward()
        if record is not None:
            record.append({"name": "E", "grads": [p.grad.detach().clone() for p in self.E.parameters()]})
        self.opt_E.step()
        if record is not None:
            record[-1]["params_after"] = [p.detach().clone() for p in self.E.parameters()]
        return {"affine_loss": loss.detach()}
