"""Shadow of colored_dSprites/utils_pxy.py (stage 1 with RGB gains; 3 + 3 codes).  See eadgan_b200/shadow."""
import argparse, itertools, math, os  # noqa: E401,F401

import numpy as np  # noqa: F401
import torch
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401
from torch.autograd import Variable  # noqa: F401
from torch.nn.utils import spectral_norm  # noqa: F401

from eadgan_b200 import affine
from eadgan_b200.shadow import _codes as K

_SPEC = (K.ZOOM1, K.SHIFT, K.SHIFT)
_RGB = (K.RGB1, K.RGB1, K.RGB1)


def from_latent_vector_2_affine_para_pxy(code_input_raw):
    return K.to_para(code_input_raw, _SPEC)


def from_affine_para_2_latent_vector_pxy(affine_para):
    return K.to_code(affine_para, _SPEC)


def from_latent_vector_2_color_para_pxy(code_input_raw):
    """gains 1 + 0.1 c"""
    return K.to_para(code_input_raw, _RGB)


def from_color_para_2_latent_vector_pxy(affine_color_para):
    return K.to_code(affine_color_para, _RGB)


def get_matrix_pxy(code_input_raw):
    """diag(p, p, 1) @ T(x, y)"""
    return K.full3(affine.pxy_matrix23(code_input_raw))


get_matrix_pxy_align_pos_size = get_matrix_pxy     # the same composition


def get_matrix_pxy_align(code_input_raw):
    """the translation alone"""
    para = from_latent_vector_2_affine_para_pxy(code_input_raw)
    m = torch.eye(3, device=para.device, dtype=para.dtype).repeat(para.shape[0], 1, 1)
    return torch.cat((m[:, :, :2], torch.stack((para[:, 1], para[:, 2], torch.ones_like(para[:, 0])), dim=1).unsqueeze(2)), dim=2)


def get_enlarge_matrix(code_input_raw):
    m = torch.eye(3, device=code_input_raw.device).repeat(code_input_raw.shape[0], 1, 1)
    return m * m.new_tensor([0.6, 0.6, 1.0]).view(1, 3, 1)


def affine_regularzier_pxy(real_code, trans_code):
    """relative zoom / shift of the first three codes, ratio of the RGB gains for the rest"""
    return affine.pxy_relative_code(real_code, trans_code).float()
