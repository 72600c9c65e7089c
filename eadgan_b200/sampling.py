"""Device-side sampling of the latent draws of a training iteration.

The reference draws z ~ N(0,1), code ~ U(-1,1) and the class labels with NumPy on the host every iteration and
copies them to the device (celebA/EAD-GAN_celebA.py:308-318; dSprites/rp.py:389-396,424-434;
colored_dSprites/rp_color.py:372-378,409-414,447-453).  ``DeviceSampler`` draws them on the GPU with the
counter-based generator of csrc/sample.cu (Philox4x32-10): no host RNG, no H2D copy, capturable in the whole-step
CUDA graph (the iteration number lives in a device counter the graph advances), and reproducible on the host word
for word (oracle/philox_ref.py, tests/test_sampling_gpu.py).  Under data parallelism every rank passes its row
offset: the union of the shards is exactly what a single device draws for the global batch.

This replaces the *generator* (NumPy's Mersenne Twister cannot be reproduced by a counter-based device stream);
the distributions and the order / shapes of the draws are the reference's."""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import call, ptr, stream

UNIFORM, NORMAL, RANDINT = 0, 1, 2


class DeviceSampler:
    def __init__(self, seed, device, row0=0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.device = torch.device(device)
        self.row0 = int(row0)
        self.step = torch.zeros(1, device=self.device, dtype=torch.int64)   # iteration number, on the device
        self._host_step = 0

    def _draw(self, kind, stream_id, rows, cols, lo=0.0, hi=1.0, n=0):
        out = torch.empty((rows, cols), device=self.device, dtype=torch.int64 if kind == RANDINT else torch.float32)
        call("eadgan_philox", kind, C.c_ulonglong(self.seed), ptr(self.step), 0, stream_id, self.row0, rows, cols,
             float(lo), float(hi), int(n), ptr(out), stream())
        return out

    def uniform(self, stream_id, rows, cols, lo, hi):
        return self._draw(UNIFORM, stream_id, rows, cols, lo, hi)

    def normal(self, stream_id, rows, cols):
        return self._draw(NORMAL, stream_id, rows, cols)

    def randint(self, stream_id, rows, n):
        return self._draw(RANDINT, stream_id, rows, 1, n=n).view(rows)

    def advance(self):
        """next iteration (a device-side increment: capturable)"""
        call("eadgan_adam_advance", ptr(self.step), stream())
        self._host_step += 1

    # ---- the draws of each training script, in the reference's order -------------------------------------
    def celeba(self, batch, latent=200, code_dim=8, n_classes=10):
        """celebA/EAD-GAN_celebA.py:308-317 -> (z [B,200], code [B,8], labels [B] int64)"""
        return (self.normal(0, batch, latent), self.uniform(1, batch, code_dim, -1.0, 1.0),
                self.randint(2, batch, n_classes))

    def dsprites(self, batch, code_dim=4, n_classes=3):
        """dSprites/rp.py:389-394 (phase D) and :424-431 (phase info) -> code_d, labels_d, code_info, labels_info"""
        return (self.uniform(1, batch, code_dim, -1.0, 1.0), self.randint(2, batch, n_classes),
                self.uniform(3, batch, code_dim, -1.0, 1.0), self.randint(4, batch, n_classes))

    def colored(self, batch, code_dim=7, n_classes=3):
        """colored_dSprites/rp_color.py:372-378 (RGB gains U(0.5,1), float64 in the reference), then the two draws"""
        gains = self.uniform(5, batch, 3, 0.5, 1.0).double().view(batch, 3, 1, 1)
        return (gains,) + self.dsprites(batch, code_dim, n_classes)


class SampledStep:
    """A training step whose latent draws come from a DeviceSampler: ``__call__(images)`` -> losses.  This is the
    end-to-end form of an iteration (the only host input is the image batch); eadgan_b200.graph.GraphedStep captures
    it whole, sampling kernels and iteration counter included."""

    def __init__(self, step, sampler, kind):
        self.step, self.sampler, self.kind = step, sampler, kind

    def optimizers(self):
        return self.step.optimizers()

    def __call__(self, images):
        B = images.shape[0]
        draws = getattr(self.sampler, self.kind)(B)
        self.sampler.advance()
        if self.kind == "celeba":
            return self.step(images, *draws)
        return self.step(images, *draws)
