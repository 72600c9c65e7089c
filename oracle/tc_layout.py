"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU (numpy) restatement of the operand layouts of the wide tcgen05 conv kernels (eadgan_b200/csrc/tc_conv.cu),
pinned against torch's conv2d / conv_transpose2d in tests/test_cpu.py:

  * FPROP (Conv2d k4 s2 p1): GEMM K = 16 c ordered (a, b, dy, dx, ci) with ky = 2a + dy, kx = 2b + dx; tap (a, b, dy)
    of output pixel (oy, ox) is the 2c-wide run [dx*c + ci] of the space-to-depth view Xp[2(oy+a)+dy][2(ox+b)+dx]
    of the halo-padded map (map_big_s2d); weights Wf[ko][tap16 * c + ci], tap16 = ((a*2+b)*2+dy)*2+dx
    (pack_w_fprop_kernel).
  * DGRAD (ConvTranspose2d k4 s2 p1 = Conv2d input gradient): four output-parity sub-GEMMs, K = 4 k ordered (ty, tx, ki);
    tap (ty, tx) of parity (py, px) reads the small map shifted by dy(py,ty), dx(px,tx) in {-1, 0, 1} and the weight
    tap ky(py,ty), kx(px,tx); Wd[(py*2+px)*c + co][(ty*2+tx)*k + ki]  (pack_w_dgrad_kernel).
  * WGRAD: dw[ko][tap16*c + ci] accumulated over pixels, then permuted to dw[ko][ci][ky][kx] (wgrad_reduce_kernel).
"""
import numpy as np


def _pad(x):
    return np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))


def fprop_patches(x):
    """x [n, c, h, w] -> A [n, p, q, 16c] in the kernel's K order (a, b, dy, dx, ci)."""
    n, c, h, w = x.shape
    p, q = h // 2, w // 2
    xp = _pad(x)
    cols = []
    for a in range(2):
        for b in range(2):
            for dy in range(2):
                for dx in range(2):
                    ky, kx = 2 * a + dy, 2 * b + dx
                    cols.append(xp[:, :, ky:ky + 2 * p:2, kx:kx + 2 * q:2].transpose(0, 2, 3, 1))   # [n, p, q, c]
    return np.concatenate(cols, axis=3)


def pack_fprop(w):
    """w [k, c, 4, 4] -> Wf [k, 16c], column tap16*c + ci."""
    k, c = w.shape[:2]
    out = np.zeros((k, 16, c), dtype=w.dtype)
    for tap in range(16):
        dx, dy, b, a = tap & 1, (tap >> 1) & 1, (tap >> 2) & 1, tap >> 3
        out[:, tap, :] = w[:, :, 2 * a + dy, 2 * b + dx]
    return out.reshape(k, 16 * c)


def fprop(x, w):
    return np.einsum("npqe,ke->nkpq", fprop_patches(x), pack_fprop(w))


def wgrad(x, dy):
    """dw [k, c, 4, 4]: GEMM over pixels in the (a, b, dy, dx, ci) column order, then the reduce kernel's permutation."""
    c = x.shape[1]
    acc = np.einsum("nkpq,npqe->ke", dy, fprop_patches(x)).reshape(-1, 16, c)
    dw = np.zeros((acc.shape[0], c, 4, 4), dtype=acc.dtype)
    for tap in range(16):
        dx, dyy, b, a = tap & 1, (tap >> 1) & 1, (tap >> 2) & 1, tap >> 3
        dw[:, :, 2 * a + dyy, 2 * b + dx] = acc[:, tap, :]
    return dw


def _tap(par, t):
    """(input shift, kernel index) of tap t in {0, 1} of output parity par in {0, 1} along one axis."""
    if par == 0:
        return (0, 1) if t == 0 else (-1, 3)
    return (1, 0) if t == 0 else (0, 2)


def pack_dgrad(w):
    """w [k, c, 4, 4] (ConvTranspose layout [Cin = k, Cout = c]) -> Wd [4c, 4k]."""
    k, c = w.shape[:2]
    out = np.zeros((4, c, 4, k), dtype=w.dtype)
    for py in range(2):
        for px in range(2):
            for ty in range(2):
                for tx in range(2):
                    out[py * 2 + px, :, ty * 2 + tx, :] = w[:, :, _tap(py, ty)[1], _tap(px, tx)[1]].T
    return out.reshape(4 * c, 4 * k)


def dgrad(y, w):
    """ConvTranspose2d(k, c, 4, 2, 1) forward of the small map y [n, k, p, q] as four parity sub-GEMMs."""
    n, k, p, q = y.shape
    c = w.shape[1]
    yp = _pad(y)
    wd = pack_dgrad(w).reshape(4, c, 4 * k)
    out = np.zeros((n, c, 2 * p, 2 * q), dtype=y.dtype)
    for py in range(2):
        for px in range(2):
            cols = []
            for ty in range(2):
                for tx in range(2):
                    dy, dx = _tap(py, ty)[0], _tap(px, tx)[0]
                    cols.append(yp[:, :, 1 + dy:1 + dy + p, 1 + dx:1 + dx + q].transpose(0, 2, 3, 1))   # [n, p, q, k]
            a = np.concatenate(cols, axis=3)                                                              # K = (ty, tx, ki)
            out[:, :, py::2, px::2] = np.einsum("npqe,ce->ncpq", a, wd[py * 2 + px])
    return out
