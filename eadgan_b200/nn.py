"""Drop-in replacements for the ``torch.nn`` modules EAD-GAN's scripts construct.

Same constructor signatures, parameter / buffer names, registration order and
``state_dict`` layout as the stock modules (each class subclasses its torch namesake
for the *host-side* bookkeeping only); ``forward`` dispatches to the sm_100a kernels
through eadgan_b200.functional.  Class names keep the substrings ``Conv`` /
``BatchNorm`` that ``weights_init_normal`` keys on (MNIST/EAD-GAN_rpqmnxy.py:54-60).

Reference constructor call sites: celebA/EAD-GAN_celebA.py:75-92,109-122,161-164;
dSprites/rp.py:65-80,94-110,128-146,164-183,249-251; MNIST/EAD-GAN_rpqmnxy.py:77-91,
105-124,141-163,195-198.
"""
from __future__ import annotations

import importlib

import torch
import torch.nn as tnn

_tsn = importlib.import_module("torch.nn.utils.spectral_norm")  # the module, not the function

from . import functional as Fn
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH

# keep handles on the stock classes: patch() rebinds the names on torch.nn
_TorchSequential = tnn.Sequential
_TorchConv2d = tnn.Conv2d
_TorchConvTranspose2d = tnn.ConvTranspose2d
_TorchLinear = tnn.Linear
_TorchBatchNorm2d = tnn.BatchNorm2d
_TorchLeakyReLU = tnn.LeakyReLU
_TorchReLU = tnn.ReLU
_TorchTanh = tnn.Tanh
_TorchSigmoid = tnn.Sigmoid
_TorchSoftmax = tnn.Softmax
_TorchUpsample = tnn.Upsample
_TorchBCELoss = tnn.BCELoss
_TorchMSELoss = tnn.MSELoss
_TorchCrossEntropyLoss = tnn.CrossEntropyLoss


def _single(v, what):
    if isinstance(v, (tuple, list)):
        if len(set(v)) != 1:
            raise RuntimeError(f"eadgan_b200: anisotropic {what} {v} is not on the reference's hot path")
        return int(v[0])
    return int(v)


class _InvalidateOnLoad:
    """mix-in: loading a state_dict rewrites parameters in place -> drop packed-weight caches"""

    def _load_from_state_dict(self, *args, **kwargs):
        from . import _lib
        _lib.bump_weights_epoch()
        self.__dict__.pop("_eadgan_sn_queue", None)   # prefetched spectral-norm results belong to the old state
        return super()._load_from_state_dict(*args, **kwargs)


class Conv2d(_InvalidateOnLoad, _TorchConv2d):
    def forward(self, x, act=(ACT_NONE, 0.0)):
        if self.groups != 1 or _single(self.dilation, "dilation") != 1 or self.padding_mode != "zeros":
            raise RuntimeError("eadgan_b200.Conv2d: groups/dilation/padding_mode variants are unsupported")
        return Fn.conv2d(x, self.weight, self.bias, _single(self.stride, "stride"),
                         _single(self.padding, "padding"), act[0], act[1])


class ConvTranspose2d(_InvalidateOnLoad, _TorchConvTranspose2d):
    def forward(self, x, output_size=None, act=(ACT_NONE, 0.0)):
        if (self.groups != 1 or _single(self.dilation, "dilation") != 1 or output_size is not None
                or _single(self.output_padding, "output_padding") != 0):
            raise RuntimeError("eadgan_b200.ConvTranspose2d: groups/dilation/output_padding are unsupported")
        return Fn.conv_transpose2d(x, self.weight, self.bias, _single(self.stride, "stride"),
                                   _single(self.padding, "padding"), act[0], act[1])


class Linear(_TorchLinear):
    def forward(self, x, act=(ACT_NONE, 0.0)):
        return Fn.linear(x, self.weight, self.bias, act[0], act[1])


class BatchNorm2d(_TorchBatchNorm2d):
    def forward(self, x, act=(ACT_NONE, 0.0)):
        self._check_input_dim(x)
        if not (self.affine and self.track_running_stats) or self.momentum is None:
            raise RuntimeError("eadgan_b200.BatchNorm2d: only affine, running-stat, fixed-momentum BN is supported")
        if self.training:
            self.num_batches_tracked.add_(1)  # torch/nn/modules/batchnorm.py::_BatchNorm.forward
        return Fn.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var, self.training,
                             self.momentum, self.eps, act[0], act[1])


class _Act:
    """mix-in: (kind, slope) of a pointwise activation module."""

    def act(self):
        raise NotImplementedError


class LeakyReLU(_TorchLeakyReLU, _Act):
    def act(self):
        return (ACT_LRELU, float(self.negative_slope))

    def forward(self, x):
        return Fn.activation(x, ACT_LRELU, self.negative_slope, self.inplace)


class ReLU(_TorchReLU, _Act):
    def act(self):
        return (ACT_RELU, 0.0)

    def forward(self, x):
        return Fn.activation(x, ACT_RELU, 0.0, self.inplace)


class Tanh(_TorchTanh, _Act):
    def act(self):
        return (ACT_TANH, 0.0)

    def forward(self, x):
        return Fn.activation(x, ACT_TANH)


class Sigmoid(_TorchSigmoid, _Act):
    def act(self):
        return (ACT_SIGMOID, 0.0)

    def forward(self, x):
        return Fn.activation(x, ACT_SIGMOID)


class Softmax(_TorchSoftmax):
    def forward(self, x):
        dim = self.dim
        if dim is None:  # implicit-dim rule of F.softmax: 1 for 2-D inputs
            dim = 0 if x.dim() in (0, 1, 3) else 1
        if x.dim() != 2 or dim not in (1, -1):
            raise RuntimeError("eadgan_b200.Softmax: only row softmax of [N, C] inputs is supported")
        return Fn.softmax(x)


class Upsample(_TorchUpsample):
    def forward(self, x):
        sf = self.scale_factor
        sf = sf[0] if isinstance(sf, (tuple, list)) else sf
        if self.mode != "nearest" or self.size is not None or float(sf) != 2.0:
            raise RuntimeError("eadgan_b200.Upsample: only nearest, scale_factor=2 is supported")
        return Fn.upsample2x(x)


_FUSABLE = (Conv2d, ConvTranspose2d, Linear, BatchNorm2d)


def _plain(m):
    return not (m._forward_hooks or m._forward_pre_hooks or m._backward_hooks)


class Sequential(_TorchSequential):
    """nn.Sequential whose forward fuses [conv|convT|linear|bn] + activation pairs into one
    kernel epilogue, and hands whole conv stacks to the bf16 tcgen05 chain executor when
    EADGAN_PRECISION=bf16 (eadgan_b200.chain)."""

    def forward(self, x):
        from . import chain
        out = chain.try_run(self, x)
        if out is not None:
            return out
        mods = list(self._modules.values())
        log = None
        if chain.gate_log is not None:     # tests only (tests/gates.py): record the ReLU / LeakyReLU gates of this forward
            log = []
            chain.gate_log.append((self, log))
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if isinstance(m, _FUSABLE) and isinstance(nxt, _Act) and _plain(nxt):
                x = m(x, act=nxt.act())
                gated = nxt
                i += 2
            else:
                x = m(x)
                gated = m if isinstance(m, _Act) else None
                i += 1
            if log is not None and gated is not None and gated.act()[0] in (ACT_RELU, ACT_LRELU):
                log.append(x.detach() > 0)
        return x


# ------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------
def _mean_only(mod):
    if mod.reduction != "mean" or getattr(mod, "weight", None) is not None:
        raise RuntimeError(f"eadgan_b200.{type(mod).__name__}: only reduction='mean' without weights is supported")


class BCELoss(_TorchBCELoss):
    def forward(self, input, target):
        _mean_only(self)
        return Fn.bce_loss(input, target)


class MSELoss(_TorchMSELoss):
    def forward(self, input, target):
        _mean_only(self)
        return Fn.mse_loss(input, target)


class CrossEntropyLoss(_TorchCrossEntropyLoss):
    def forward(self, input, target):
        _mean_only(self)
        if self.ignore_index != -100 or self.label_smoothing != 0.0:
            raise RuntimeError("eadgan_b200.CrossEntropyLoss: ignore_index / label_smoothing are unsupported")
        return Fn.cross_entropy(input, target)


# ------------------------------------------------------------------------------
# legacy spectral_norm
# ------------------------------------------------------------------------------
class SpectralNorm(_tsn.SpectralNorm):
    """torch's legacy SpectralNorm hook object with compute_weight routed to the fused
    power-iteration kernels.  Buffer / parameter names and the state_dict hooks are
    torch's own (weight_orig, weight_u, weight_v, version metadata)."""

    def compute_weight(self, module, do_power_iteration):
        weight = getattr(module, self.name + "_orig")
        u = getattr(module, self.name + "_u")
        v = getattr(module, self.name + "_v")
        if self.dim != 0:
            raise RuntimeError("eadgan_b200.spectral_norm: dim != 0 (ConvTranspose) is not used by the reference")
        if self.n_power_iterations != 1:
            raise RuntimeError("eadgan_b200.spectral_norm: n_power_iterations must be 1")
        w_sn, sigma = Fn.spectral_norm_weight(weight, u, v, do_power_iteration, self.eps)
        module._eadgan_sigma = sigma
        # (weight_orig, sigma, the tensor about to become module.weight): lets the bf16 chain pack weight_orig
        # once per optimiser step and apply 1/sigma in the conv epilogue (eadgan_b200.chain.try_run)
        module.__dict__["_eadgan_sn_src"] = (weight, sigma, w_sn, Fn.sn_skip_scale)
        return w_sn


def spectral_norm(module, name="weight", n_power_iterations=1, eps=1e-12, dim=None):
    """Same registration as torch.nn.utils.spectral_norm (so seeded u/v init and the
    state_dict layout are identical); only the per-forward computation differs."""
    if dim is None:
        dim = 1 if isinstance(module, (tnn.ConvTranspose1d, _TorchConvTranspose2d, tnn.ConvTranspose3d)) else 0
    fn = _tsn.SpectralNorm.apply(module, name, n_power_iterations, dim, eps)
    fn.__class__ = SpectralNorm
    return module
