"""Shadow of colored_dSprites/utils_rp_color.py (4 affine codes + 3 colour codes).  See eadgan_b200/shadow."""
import argparse, itertools, math, os  # noqa: E401,F401

import numpy as np  # noqa: F401
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401
from torch.autograd import Variable  # noqa: F401
from torch.nn.utils import spectral_norm  # noqa: F401

from eadgan_b200 import affine
from eadgan_b200.shadow import _codes as K

_SPEC = (K.THETA, K.ZOOM2, K.SHIFT, K.SHIFT)
_RGB = (K.RGB5, K.RGB5, K.RGB5)


def from_latent_vector_2_affine_para(code_input_raw):
    """:24-35"""
    return K.to_para(code_input_raw, _SPEC)


def from_latent_vector_2_color_para(code_input_raw):
    """:38-46  gains 1 + 0.5 c"""
    return K.to_para(code_input_raw, _RGB)


def from_affine_para_2_latent_vector(affine_color_para):
    """:49-61"""
    return K.to_code(affine_color_para, _SPEC)


def from_color_para_2_latent_vector(affine_color_para):
    """:64-73"""
    return K.to_code(affine_color_para, _RGB)


def get_matrix(code_input_raw):
    """:76-96  R(theta) @ diag(p, p, 1) @ T(x, y)"""
    return K.full3(affine.dsprites_matrix23(code_input_raw))


def affine_color_regularzier(real_code, trans_code):
    """:99-139"""
    return affine.colored_relative_code(real_code, trans_code).float()
