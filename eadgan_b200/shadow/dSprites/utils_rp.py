"""Shadow of dSprites/utils_rp.py (rotation, isotropic zoom, translation; 4 codes).  See eadgan_b200/shadow."""
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


def from_latent_vector_2_affine_para_D(code_input_raw):
    """:23-35"""
    return K.to_para(code_input_raw, _SPEC)


def from_latent_vector_2_affine_para(code_input_raw):
    """:62-74 (the same map)"""
    return K.to_para(code_input_raw, _SPEC)


def from_affine_para_2_latent_vector(affine_color_para):
    """:77-91"""
    return K.to_code(affine_color_para, _SPEC)


def get_matrix(code_input_raw):
    """:94-115  R(theta) @ diag(p, p, 1) @ T(x, y) as [B, 3, 3]"""
    return K.full3(affine.dsprites_matrix23(code_input_raw))


get_matrix_D = get_matrix      # :38-59, the same composition


def affine_regularzier(real_code, trans_code):
    """:117-147"""
    return affine.dsprites_relative_code(real_code, trans_code).float()
