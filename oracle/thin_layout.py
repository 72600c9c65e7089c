"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU (numpy) restatement of the DATA LAYOUT and index algebra of the "thin" image-layer kernels
(eadgan_b200/csrc/tc_conv.cu, section "Thin image layers"), so that the layout claims made in include/eadgan.h and
DESIGN.md are pinned on the CPU against torch's own conv2d / conv_transpose2d (tests/test_cpu.py):

  * row-expanded image buffer  R[n][oy][X][ky][c] = Xpad[n][2 oy + ky][X][c]   (thin_expand_kernel)
  * thin fprop: the patch of output pixel (oy, ox) is R[n, oy, 2 ox : 2 ox + 4] flattened = 64 elements in
    (kx, ky, c) order; weights packed as Wf[ko][kx*16 + ky*4 + c]            (thin_pack_kernel, direction 0)
  * thin wgrad: dw[ko][c][ky][kx] = sum_pixels dy[pixel][ko] * patch[pixel][kx*16 + ky*4 + c]
  * thin dgrad (ConvTranspose2d k4 s2 p1 onto the image): Z[pixel][ky*16 + kx*4 + c] = y[pixel][:] . Wt, then the
    col2im ownership rule of tc_thin_dgrad_kernel: input pixel (m, j) owns output pixels (2m + a, 2j + b) and
        out[2m  ][2j+b] = Xr_m[ky 1][b] + Xr_{m-1}[ky 3][b]      Xr[ky][0] = Z[ky][kx 1] + Z_{j-1}[ky][kx 3]
        out[2m+1][2j+b] = Xr_m[ky 2][b] + Xr_{m+1}[ky 0][b]      Xr[ky][1] = Z[ky][kx 2] + Z_{j+1}[ky][kx 0]
"""
import numpy as np


def expand(x):
    """x [n, c<=4, h, w] -> R [n, h/2, w+2, 4(ky), 4(c)] (zero halo, zero padding channels)."""
    n, c, h, w = x.shape
    xp = np.zeros((n, 4, h + 2, w + 2), dtype=x.dtype)
    xp[:, :c, 1:-1, 1:-1] = x
    r = np.zeros((n, h // 2, w + 2, 4, 4), dtype=x.dtype)
    for ky in range(4):
        r[:, :, :, ky, :] = xp[:, :, ky:ky + h:2, :].transpose(0, 2, 3, 1)
    return r


def patches(r):
    """R -> [n, p, q, 64]: element kx*16 + ky*4 + c of pixel (oy, ox) -- 64 CONTIGUOUS elements of R starting at X = 2 ox
    (the overlapping-window TMA box of map_thin: 64-element extent, 32-element stride)."""
    n, p, wp, _, _ = r.shape
    q = (wp - 2) // 2
    flat = r.reshape(n, p, wp * 16)
    return np.stack([flat[:, :, 32 * ox:32 * ox + 64] for ox in range(q)], axis=2)


def pack_fprop(w):
    """w [k, c, 4, 4] -> Wf [k, 64] with column kx*16 + ky*4 + c (zero for c >= c_real)."""
    k, c = w.shape[:2]
    out = np.zeros((k, 4, 4, 4), dtype=w.dtype)          # [ko][kx][ky][c]
    out[:, :, :, :c] = w.transpose(0, 3, 2, 1)
    return out.reshape(k, 64)


def pack_dgrad(w):
    """w [k, c, 4, 4] -> Wt [64, k] with row ky*16 + kx*4 + c."""
    k, c = w.shape[:2]
    out = np.zeros((4, 4, 4, k), dtype=w.dtype)          # [ky][kx][c][ki]
    out[:, :, :c, :] = w.transpose(2, 3, 1, 0)
    return out.reshape(64, k)


def fprop(x, w):
    """Conv2d(c, k, 4, 2, 1) forward as ONE K = 64 GEMM per pixel."""
    return np.einsum("npqe,ke->nkpq", patches(expand(x)), pack_fprop(w))


def wgrad(x, dy):
    """dw [k, c, 4, 4] from the image x and the small-map gradient dy [n, k, p, q]."""
    c = x.shape[1]
    acc = np.einsum("nkpq,npqe->ke", dy, patches(expand(x))).reshape(-1, 4, 4, 4)   # [ko][kx][ky][c]
    return acc.transpose(0, 3, 2, 1)[:, :c]


def dgrad(y, w):
    """ConvTranspose2d(k, c, 4, 2, 1) forward: GEMM over the small map's pixels + col2im with the kernel's ownership rule."""
    n, k, p, q = y.shape
    c = w.shape[1]
    z = np.einsum("nkpq,ek->npqe", y, pack_dgrad(w)).reshape(n, p, q, 4, 4, 4)        # [.., ky, kx, c]
    zp = np.zeros((n, p + 2, q + 2, 4, 4, 4), dtype=z.dtype)                          # zero halo rows / columns
    zp[:, 1:-1, 1:-1] = z
    # x direction: Xr[ky][b]
    xr = np.zeros((n, p + 2, q, 4, 2, 4), dtype=z.dtype)
    xr[..., 0, :] = zp[:, :, 1:-1, :, 1, :] + zp[:, :, :-2, :, 3, :]                  # own kx 1 + left neighbour's kx 3
    xr[..., 1, :] = zp[:, :, 1:-1, :, 2, :] + zp[:, :, 2:, :, 0, :]                   # own kx 2 + right neighbour's kx 0
    out = np.zeros((n, c, 2 * p, 2 * q), dtype=z.dtype)
    for b in range(2):
        even = xr[:, 1:-1, :, 1, b, :c] + xr[:, :-2, :, 3, b, :c]                      # row m ky 1 + row m-1 ky 3
        odd = xr[:, 1:-1, :, 2, b, :c] + xr[:, 2:, :, 0, b, :c]                        # row m ky 2 + row m+1 ky 0
        out[:, :, 0::2, b::2] = even.transpose(0, 3, 1, 2)
        out[:, :, 1::2, b::2] = odd.transpose(0, 3, 1, 2)
    return out
