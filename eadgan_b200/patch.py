"""patch()/unpatch(): rebind the torch names EAD-GAN's scripts use to the eadgan_b200
replacements, so the UNMODIFIED reference scripts run on the sm_100a kernels
(``python -m eadgan_b200.run <script.py> [args]``).

Patched names: torch.nn.{Sequential, Conv2d, ConvTranspose2d, Linear, BatchNorm2d, LeakyReLU,
ReLU, Tanh, Sigmoid, Softmax, Upsample, BCELoss, MSELoss, CrossEntropyLoss},
torch.nn.utils.spectral_norm, torch.nn.functional.{sigmoid, softmax, affine_grid, grid_sample},
torch.optim.Adam, torch.inverse (batched 3x3 on CUDA: closed form, no host synchronisation).
"""
from __future__ import annotations

import torch
import torch.nn.functional as TF

from . import functional as Fn
from . import nn as enn
from . import optim as eoptim
from ._lib import ACT_SIGMOID

_NN_NAMES = ["Sequential", "Conv2d", "ConvTranspose2d", "Linear", "BatchNorm2d", "LeakyReLU", "ReLU", "Tanh",
             "Sigmoid", "Softmax", "Upsample", "BCELoss", "MSELoss", "CrossEntropyLoss"]
_saved = {}


def sigmoid(input):
    """F.sigmoid drop-in (celebA/EAD-GAN_celebA.py:130, dSprites/rp.py:117,155)."""
    return Fn.activation(input, ACT_SIGMOID)


def softmax(input, dim=None, _stacklevel=3, dtype=None):
    """F.softmax drop-in with the implicit-dim rule (celebA/EAD-GAN_celebA.py:132)."""
    if dim is None:
        dim = 0 if input.dim() in (0, 1, 3) else 1
    if input.dim() != 2 or dim not in (1, -1) or dtype is not None:
        raise RuntimeError("eadgan_b200 softmax: only row softmax of [N, C] inputs is supported")
    return Fn.softmax(input)


def inverse(input, *, out=None):
    """torch.inverse drop-in: batched CUDA [B, 3, 3] matrices (every use in the reference: the 3x3 affine matrices of
    utils_*.py) go through the closed-form adjugate on the device; anything else through stock torch."""
    if out is None and torch.is_tensor(input) and input.is_cuda and input.dim() == 3 and tuple(input.shape[1:]) == (3, 3):
        from . import affine
        return affine.inverse3x3(input)
    return _saved[("torch", "inverse")](input) if out is None else _saved[("torch", "inverse")](input, out=out)


def patch():
    if _saved:
        return
    for n in _NN_NAMES:
        _saved[("nn", n)] = getattr(torch.nn, n)
        setattr(torch.nn, n, getattr(enn, n))
    _saved[("utils", "spectral_norm")] = torch.nn.utils.spectral_norm
    torch.nn.utils.spectral_norm = enn.spectral_norm
    _saved[("F", "sigmoid")] = TF.sigmoid
    _saved[("F", "softmax")] = TF.softmax
    TF.sigmoid = sigmoid
    TF.softmax = softmax
    # transformation_2D.stn of every script: F.affine_grid + F.grid_sample, forward AND backward (dSprites/rp.py reaches
    # the backward through its grad-tracked frozen Encoder_pxy)
    _saved[("F", "affine_grid")] = TF.affine_grid
    _saved[("F", "grid_sample")] = TF.grid_sample
    TF.affine_grid = Fn.affine_grid
    TF.grid_sample = Fn.grid_sample
    _saved[("optim", "Adam")] = torch.optim.Adam
    torch.optim.Adam = eoptim.Adam
    _saved[("torch", "inverse")] = torch.inverse
    torch.inverse = inverse


def unpatch():
    for (where, n), v in list(_saved.items()):
        if where == "nn":
            setattr(torch.nn, n, v)
        elif where == "utils":
            torch.nn.utils.spectral_norm = v
        elif where == "F":
            setattr(TF, n, v)
        elif where == "optim":
            torch.optim.Adam = v
        elif where == "torch":
            setattr(torch, n, v)
    _saved.clear()
