"""Shadow of celebA/utils_rpqxy.py (rotation, anisotropic zoom p / q, translation; 5 codes).  See eadgan_b200/shadow."""
import argparse, itertools, math, os  # noqa: E401,F401  (the scripts star-import this module: keep the names it exposed)

import numpy as np  # noqa: F401
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401
from torch.autograd import Variable  # noqa: F401
from torch.nn.utils import spectral_norm  # noqa: F401

from eadgan_b200 import affine
from eadgan_b200.shadow import _codes as K

_SPEC = (K.THETA, K.ZOOM2, K.ZOOM2, K.SHIFT, K.SHIFT)


def from_latent_vector_2_affine_para(code_input_raw):
    """:25-38  code -> (theta, p, q, x, y)"""
    return K.to_para(code_input_raw, _SPEC)


def from_affine_para_2_latent_vector(affine_color_para):
    """:41-55"""
    return K.to_code(affine_color_para, _SPEC)


def get_matrix(code_input_raw):
    """:59-80  R(theta) @ diag(p, q, 1) @ T(x, y) as [B, 3, 3], composed on the device"""
    return affine.celeba_matrix(code_input_raw)


def affine_regularzier(real_code, trans_code):
    """:82-116  parameters of M(trans) @ inverse(M(real)) mapped back to a code: one fused kernel"""
    return affine.celeba_relative_code(real_code, trans_code).float()
