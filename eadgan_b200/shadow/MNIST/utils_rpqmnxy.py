"""Shadow of MNIST/utils_rpqmnxy.py (rotation, zoom p / q, skew m / n, translation; 7 codes; the relative parameters
are read off by a frozen MLP approximator loaded from ``rpqmnxy_approximator.pt`` in the working directory, exactly
as the reference does at import time).  See eadgan_b200/shadow."""
import numpy as np  # noqa: F401
import torch
import torch.nn as nn
from torch.autograd import Variable  # noqa: F401

from eadgan_b200 import affine
from eadgan_b200.shadow import _codes as K

cuda = True if torch.cuda.is_available() else False
FloatTensor = torch.cuda.FloatTensor if cuda else torch.FloatTensor
_SPEC = (K.THETA, K.ZOOM2, K.ZOOM2, K.SKEW, K.SKEW, K.SHIFT, K.SHIFT)


class Affine_classifier(nn.Module):
    """:11-33  6 -> 256 -> 256 -> 256 -> 256 -> 7, LeakyReLU(0.01) between"""

    def __init__(self):
        super().__init__()
        widths = [6, 256, 256, 256, 256]
        layers = []
        for i, o in zip(widths[:-1], widths[1:]):
            layers += [nn.Linear(i, o), nn.LeakyReLU()]
        self.fc_block = nn.Sequential(*layers, nn.Linear(256, 7))

    def forward(self, real_transfrom_code):
        return self.fc_block(real_transfrom_code)


PATH = "rpqmnxy_approximator.pt"
BFGS_approximator = Affine_classifier()
BFGS_approximator.cuda()
BFGS_approximator.load_state_dict(torch.load(PATH))
BFGS_approximator.eval()
BFGS_approximator.requires_grad = False      # (as in the reference: a plain attribute, the parameters stay grad-tracked)


def from_latent_vector_2_affine_para(code_input_raw):
    """:46-60"""
    return K.to_para(code_input_raw, _SPEC)


def from_affine_para_2_latent_vector(affine_para):
    """:67-84"""
    return K.to_code(affine_para, _SPEC)


def get_matrix(code_input_raw):
    """:87-114  R(theta) @ diag(p, q, 1) @ Skew(m, n) @ T(x, y)"""
    return K.full3(affine.mnist_matrix23(code_input_raw))


def affine_regularizer(real_code, trans_code):
    """:117-134  top two rows of M(trans) @ inverse(M(real)) (fused kernel) -> approximator -> code"""
    return from_affine_para_2_latent_vector(BFGS_approximator(affine.mnist_relative_rows(real_code, trans_code)))
