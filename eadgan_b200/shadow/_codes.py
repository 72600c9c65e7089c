"""latent code <-> affine parameter maps shared by the shadow modules: column i of the code is scaled and shifted as
(code * mul / div + add); any further columns stay zero, as in the reference's zero-initialised result buffers."""
import math

import torch

PI = math.pi


def to_para(code, spec):
    cols = [code[:, i] * mul / div + add for i, (mul, div, add) in enumerate(spec)]
    cols += [torch.zeros_like(code[:, 0])] * (code.shape[1] - len(spec))
    return torch.stack(cols, dim=1)


def to_code(para, spec):
    """inverse of to_para, in the reference's operation order: (para - add) / mul * div"""
    cols = [(para[:, i] - add) / mul * div if add else para[:, i] / mul * div for i, (mul, div, add) in enumerate(spec)]
    cols += [torch.zeros_like(para[:, 0])] * (para.shape[1] - len(spec))
    return torch.stack(cols, dim=1)


def full3(m23):
    """[B, 2, 3] -> [B, 3, 3] with the affine last row (0, 0, 1)"""
    last = m23.new_zeros((m23.shape[0], 1, 3))
    last[:, 0, 2] = 1.0
    return torch.cat((m23, last), dim=1)


THETA, ZOOM2, SHIFT, ZOOM1, SKEW, RGB5, RGB1 = ((PI, 9.0, 0.0), (0.2, 1.0, 1.0), (0.1, 1.0, 0.0), (0.1, 1.0, 1.0),
                                               (0.2, 1.0, 0.0), (0.5, 1.0, 1.0), (0.1, 1.0, 1.0))
