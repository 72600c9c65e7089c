"""Affine-code glue of the training step: latent code -> 3x3 matrix -> STN warp, and the
closed-form recovery of relative affine parameters (reference: celebA/utils_rpqxy.py:25-116,
dSprites/utils_rp.py:23-147, dSprites/utils_pxy.py:24-126).

SURVEY.md section 8(f) ranks 1-2.  The reference builds ``torch.eye(3)`` on the CPU and assigns
CUDA slices into it, i.e. 8-9 synchronising D2H copies per call.  Here the
algebra is restated in closed form (``*_torch`` functions: plain differentiable torch ops, no host round
trip; they run on any device and are what the CPU tests pin to the oracle), and on CUDA tensors the public
functions dispatch to ONE fused kernel each (csrc/glue.cu): the relative-code recovery evaluated in
forward-mode dual numbers with a stored Jacobian (backward = J^T g), and the spatial transformer
(affine_grid + bilinear grid_sample) as a single forward kernel.
"""
from __future__ import annotations

import math

import torch

from . import functional as Fn


def _fused(*ts):
    return all(t.is_cuda and t.dtype == torch.float32 for t in ts)


def _require_fused(who, *ts):
    """the product entry points run the fused kernels only: no CPU / other-dtype fallback (the ``*_torch``
    restatements are for algebra checks and fp64 referees in the tests)"""
    if not _fused(*ts):
        raise RuntimeError(f"eadgan_b200.affine.{who}: CUDA float32 tensors required (no CPU fallback); "
                           f"use {who}_torch for a reference evaluation on other tensors")


def _mat3(rows):
    """rows: 3x3 nested list of [B] tensors / python floats -> [B,3,3]."""
    ref = next(e for r in rows for e in r if torch.is_tensor(e))
    out = []
    for r in rows:
        out.append(torch.stack([e if torch.is_tensor(e) else torch.full_like(ref, float(e)) for e in r], dim=1))
    return torch.stack(out, dim=1)


def compose_rzt(theta, p, q, x, y):
    """R(theta) @ diag(p,q,1) @ T(x,y) written out (celebA/utils_rpqxy.py:64-77)."""
    c, s = torch.cos(theta), torch.sin(theta)
    return _mat3([[c * p, -s * q, c * p * x - s * q * y],
                  [s * p, c * q, s * p * x + c * q * y],
                  [0.0, 0.0, 1.0]])


def celeba_matrix(code5):
    return compose_rzt(code5[:, 0] * (math.pi / 9), code5[:, 1] * 0.2 + 1, code5[:, 2] * 0.2 + 1,
                       code5[:, 3] * 0.1, code5[:, 4] * 0.1)


def _rzt_parts(code5):
    """(a, b, c, d, tx, ty) of [[a, b, tx], [c, d, ty], [0, 0, 1]] = R(theta) diag(p, q, 1) T(x, y)."""
    theta, p, q = code5[:, 0] * (math.pi / 9), code5[:, 1] * 0.2 + 1, code5[:, 2] * 0.2 + 1
    x, y = code5[:, 3] * 0.1, code5[:, 4] * 0.1
    c, s = torch.cos(theta), torch.sin(theta)
    a, b, cc, d = c * p, -s * q, s * p, c * q
    return a, b, cc, d, a * x + b * y, cc * x + d * y


def celeba_relative_code(real_code, trans_code):
    """affine_regularzier of celebA/utils_rpqxy.py:82-116 (fused kernel on CUDA tensors)."""
    _require_fused("celeba_relative_code", real_code, trans_code)
    return Fn.relative_code(real_code, trans_code, Fn.REL_CELEBA)


def celeba_relative_code_torch(real_code, trans_code):
    """affine_regularzier of celebA/utils_rpqxy.py:82-116.  rel = M(trans) @ inverse(M(real)); both are
    affine (last row 0 0 1), so the inverse is written in closed form -- torch.inverse / linalg.inv check
    for singularity on the HOST (a device->host sync in the middle of every step)."""
    a1, b1, c1, d1, x1, y1 = _rzt_parts(real_code[:, :5])
    a2, b2, c2, d2, x2, y2 = _rzt_parts(trans_code[:, :5])
    det = a1 * d1 - b1 * c1
    ia, ib, ic, id_ = d1 / det, -b1 / det, -c1 / det, a1 / det          # inverse of the 2x2 block
    itx, ity = -(ia * x1 + ib * y1), -(ic * x1 + id_ * y1)               # inverse translation
    a, b = a2 * ia + b2 * ic, a2 * ib + b2 * id_
    c, d = c2 * ia + d2 * ic, c2 * ib + d2 * id_
    tx, ty = a2 * itx + b2 * ity + x2, c2 * itx + d2 * ity + y2
    th = 0.5 * torch.atan(2 * (a * c - b * d) / (a * a + d * d - b * b - c * c))
    ct, st = torch.cos(th), torch.sin(th)
    p = a * ct + c * st
    q = -b * st + d * ct
    x = (tx * ct + ty * st) / p
    y = (ty * ct - tx * st) / q
    return torch.stack((th * (9 / math.pi), (p - 1) / 0.2, (q - 1) / 0.2, x / 0.1, y / 0.1), dim=1)


def stn(img, theta23, padding_mode="border"):
    """transformation_2D.stn (celebA/EAD-GAN_celebA.py:149-153).  When no gradient is requested through it (every
    consumed gradient of the training steps: the warped images are constants) this is ONE fused forward kernel;
    otherwise the affine_grid and grid_sample kernels with their backward passes (csrc/glue.cu) -- never stock ATen."""
    if not _fused(img, theta23):
        raise RuntimeError("eadgan_b200.affine.stn: CUDA float32 tensors required (no CPU fallback)")
    if padding_mode not in ("border", "zeros"):
        raise RuntimeError("eadgan_b200.affine.stn: padding_mode must be 'border' or 'zeros'")
    needs_grad = torch.is_grad_enabled() and (img.requires_grad or theta23.requires_grad)
    if not needs_grad:
        return Fn.stn_fwd(img, theta23, border=padding_mode == "border")
    return Fn.grid_sample(img, Fn.affine_grid(theta23, tuple(img.shape)), padding_mode=padding_mode)


# ---- dSprites (dSprites/utils_pxy.py, dSprites/utils_rp.py) -----------------------------------------------
def dsprites_align_inverse(code3):
    """inverse(get_matrix_pxy_align(code))[:, 0:2]  (dSprites/utils_pxy.py:69-87, rp.py:374-377): the align
    matrix is a pure translation by (0.1*c1, 0.1*c2) -- the zoom entry c0 is NOT used -- so its inverse is the
    opposite translation."""
    one, zero = torch.ones_like(code3[:, 0]), torch.zeros_like(code3[:, 0])
    return torch.stack((torch.stack((one, zero, -(code3[:, 1] * 0.1)), dim=1),
                        torch.stack((zero, one, -(code3[:, 2] * 0.1)), dim=1)), dim=1)


def _rpt_parts(code4):
    """(a, b, c, d, tx, ty) of R(theta) diag(p, p, 1) T(x, y)  (dSprites/utils_rp.py:38-59 == :94-115)."""
    theta, p = code4[:, 0] * (math.pi / 9), code4[:, 1] * 0.2 + 1
    x, y = code4[:, 2] * 0.1, code4[:, 3] * 0.1
    c, s = torch.cos(theta), torch.sin(theta)
    a, b, cc, d = c * p, -s * p, s * p, c * p
    return a, b, cc, d, a * x + b * y, cc * x + d * y


def dsprites_matrix23(code4):
    a, b, c, d, tx, ty = _rpt_parts(code4[:, :4])
    return torch.stack((torch.stack((a, b, tx), dim=1), torch.stack((c, d, ty), dim=1)), dim=1)


def dsprites_relative_code(real_code, trans_code):
    """affine_regularzier of dSprites/utils_rp.py:117-147 (fused kernel on CUDA tensors)."""
    _require_fused("dsprites_relative_code", real_code, trans_code)
    return Fn.relative_code(real_code, trans_code, Fn.REL_DSPRITES)


def dsprites_relative_code_torch(real_code, trans_code):
    """affine_regularzier of dSprites/utils_rp.py:117-147, closed-form inverse (no host round trip)."""
    a1, b1, c1, d1, x1, y1 = _rpt_parts(real_code[:, :4])
    a2, b2, c2, d2, x2, y2 = _rpt_parts(trans_code[:, :4])
    det = a1 * d1 - b1 * c1
    ia, ib, ic, id_ = d1 / det, -b1 / det, -c1 / det, a1 / det
    itx, ity = -(ia * x1 + ib * y1), -(ic * x1 + id_ * y1)
    r00, r01 = a2 * ia + b2 * ic, a2 * ib + b2 * id_
    r10, r11 = c2 * ia + d2 * ic, c2 * ib + d2 * id_
    r02, r12 = a2 * itx + b2 * ity + x2, c2 * itx + d2 * ity + y2
    th = torch.atan((r10 - r01) / (r00 + r11))
    ct, st = torch.cos(th), torch.sin(th)
    p = 0.5 * (ct * (r00 + r11) + st * (r10 - r01))
    x = (r02 * ct + r12 * st) / p
    y = (r12 * ct - r02 * st) / p
    return torch.stack((th * (9 / math.pi), (p - 1) / 0.2, x / 0.1, y / 0.1), dim=1)


# ---- colored dSprites (colored_dSprites/utils_rp_color.py) -------------------------------------------------
def colored_relative_code(real_code, trans_code, _affine=None):
    """affine_color_regularzier of colored_dSprites/utils_rp_color.py:99-139: entries 0..3 as in
    dsprites_relative_code; entries 4..6 are the ratio of the colour gains c * 0.5 + 1 mapped back to a code."""
    aff = (_affine or dsprites_relative_code)(real_code[:, :4], trans_code[:, :4])
    rel = (trans_code[:, 4:] * 0.5 + 1) / (real_code[:, 4:] * 0.5 + 1)
    return torch.cat((aff, (rel - 1) / 0.5), dim=1)


def colored_relative_code_torch(real_code, trans_code):
    return colored_relative_code(real_code, trans_code, _affine=dsprites_relative_code_torch)


# ---- MNIST (MNIST/utils_rpqmnxy.py) ------------------------------------------------------------------------
def _rzst_parts(code7):
    """(a, b, c, d, tx, ty) of R(theta) diag(p, q, 1) Skew(m, n) T(x, y)  (MNIST/utils_rpqmnxy.py:46-60,87-114)."""
    theta, p, q = code7[:, 0] * (math.pi / 9), code7[:, 1] * 0.2 + 1, code7[:, 2] * 0.2 + 1
    m, n = code7[:, 3] * 0.2, code7[:, 4] * 0.2
    x, y = code7[:, 5] * 0.1, code7[:, 6] * 0.1
    c, s = torch.cos(theta), torch.sin(theta)
    # R Z = [[c p, -s q], [s p, c q]];  (R Z) S with S = [[1, m], [n, 1]]
    a, b = c * p - s * q * n, c * p * m - s * q
    cc, d = s * p + c * q * n, s * p * m + c * q
    return a, b, cc, d, a * x + b * y, cc * x + d * y


def mnist_matrix23(code7):
    a, b, c, d, tx, ty = _rzst_parts(code7)
    return torch.stack((torch.stack((a, b, tx), dim=1), torch.stack((c, d, ty), dim=1)), dim=1)


def mnist_relative_rows(real_code, trans_code):
    """top two rows of M(trans) @ inverse(M(real)) as [B, 6] (fused kernel on CUDA tensors)."""
    _require_fused("mnist_relative_rows", real_code, trans_code)
    return Fn.relative_code(real_code, trans_code, Fn.REL_MNIST)


def mnist_relative_rows_torch(real_code, trans_code):
    """top two rows of M(trans) @ inverse(M(real)), flattened to [B, 6] -- the approximator's input
    (MNIST/utils_rpqmnxy.py:123-129); closed-form affine inverse, no host round trip."""
    a1, b1, c1, d1, x1, y1 = _rzst_parts(real_code)
    a2, b2, c2, d2, x2, y2 = _rzst_parts(trans_code)
    det = a1 * d1 - b1 * c1
    ia, ib, ic, id_ = d1 / det, -b1 / det, -c1 / det, a1 / det
    itx, ity = -(ia * x1 + ib * y1), -(ic * x1 + id_ * y1)
    r00, r01 = a2 * ia + b2 * ic, a2 * ib + b2 * id_
    r10, r11 = c2 * ia + d2 * ic, c2 * ib + d2 * id_
    r02, r12 = a2 * itx + b2 * ity + x2, c2 * itx + d2 * ity + y2
    return torch.stack((r00, r01, r02, r10, r11, r12), dim=1)


def mnist_code_from_params(pred):
    """from_affine_para_2_latent_vector (MNIST/utils_rpqmnxy.py:64-84)."""
    return torch.stack((pred[:, 0] * (9 / math.pi), (pred[:, 1] - 1) / 0.2, (pred[:, 2] - 1) / 0.2, pred[:, 3] / 0.2,
                        pred[:, 4] / 0.2, pred[:, 5] / 0.1, pred[:, 6] / 0.1), dim=1)


# ---- stage 1 (dSprites/utils_pxy.py, colored_dSprites/utils_pxy.py) -----------------------------------------
def pxy_matrix23(code):
    """get_matrix_pxy(code)[:, 0:2] = (Z(p,p) @ T(x,y))[:, 0:2]  (dSprites/utils_pxy.py:49-66)."""
    p = code[:, 0] * 0.1 + 1
    zero = torch.zeros_like(p)
    return torch.stack((torch.stack((p, zero, p * (code[:, 1] * 0.1)), dim=1),
                        torch.stack((zero, p, p * (code[:, 2] * 0.1)), dim=1)), dim=1)


def pxy_relative_code(real_code, trans_code):
    """affine_regularzier_pxy (dSprites/utils_pxy.py:107-126, colored_dSprites/utils_pxy.py:150-175) in closed
    form: M = [[p,0,px],[0,p,py]] so M2 @ inverse(M1) has zoom p2/p1 and translation p2 (x2 - x1)."""
    p1, p2 = real_code[:, 0] * 0.1 + 1, trans_code[:, 0] * 0.1 + 1
    rp = p2 / p1
    rx = p2 * (trans_code[:, 1] * 0.1 - real_code[:, 1] * 0.1) / rp
    ry = p2 * (trans_code[:, 2] * 0.1 - real_code[:, 2] * 0.1) / rp
    out = torch.stack(((rp - 1) / 0.1, rx / 0.1, ry / 0.1), dim=1)
    if real_code.shape[1] > 3:
        rel = (trans_code[:, 3:] * 0.1 + 1) / (real_code[:, 3:] * 0.1 + 1)
        out = torch.cat((out, (rel - 1) / 0.1), dim=1)
    return out


def inverse3x3(m):
    """batched inverse of [B, 3, 3] matrices by the adjugate (rows of the inverse's transpose are cross products of
    the rows) -- what the scripts' ``torch.inverse(get_matrix*(code))`` computes (dSprites/rp.py:376,
    celebA/utils_rpqxy.py:92), without the host synchronisation torch.linalg's singularity check costs per call.
    Differentiable through ordinary autograd."""
    r0, r1, r2 = m[:, 0], m[:, 1], m[:, 2]
    c0, c1, c2 = torch.cross(r1, r2, dim=1), torch.cross(r2, r0, dim=1), torch.cross(r0, r1, dim=1)
    det = (r0 * c0).sum(dim=1, keepdim=True)
    return torch.stack((c0 / det, c1 / det, c2 / det), dim=2)
