"""-m gpu: the bf16 tcgen05 chain executor (eadgan_b200.chain) on whole conv stacks.

Parity protocol for the bf16 mode (DESIGN.md "parity protocol"):
  1. gate-insensitive stacks (LeakyReLU slope 0.9 instead of 0.1 / ReLU): every fused path of the
     chain -- epilogue bias+activation, fused BatchNorm statistics, mask-fused input gradients, tcgen05
     wgrad, SIMT edge layers -- against the bf16-rounding torch reference (tests/bf16_emul.py): L2-relative
     error <= 2e-2 and tensor-normalised max error <= 6e-2 on outputs AND every gradient (see STACK_L2);
  2. the reference's real G / D (ReLU, LeakyReLU(0.1), spectral norm): outputs likewise against the
     fp32 torch modules; gradients by cosine similarity / L2-relative error, because ANY two bf16
     evaluations of these nets disagree on ~0.1 % of the activation gates (a pre-activation within bf16
     rounding of 0), and one flipped gate moves a 128-term bias-gradient sum by ~10 % at batch 8.
"""
import os

import pytest
import torch

from conftest import rel_err

import bf16_emul

pytestmark = pytest.mark.gpu


def _cos(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def _l2(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def _build(ns, spec):
    import eadgan_b200.nn as enn
    layers = []
    for item in spec:
        kind = item[0]
        if kind in ("conv", "snconv"):
            m = ns.Conv2d(*item[1:])
            if kind == "snconv":
                m = (enn.spectral_norm if ns is enn else torch.nn.utils.spectral_norm)(m)
            layers.append(m)
        elif kind == "convT":
            layers.append(ns.ConvTranspose2d(*item[1:]))
        elif kind == "bn":
            layers.append(ns.BatchNorm2d(item[1]))
        elif kind == "lrelu":
            layers.append(ns.LeakyReLU(item[1], inplace=True))
        elif kind == "relu":
            layers.append(ns.ReLU())
        elif kind == "tanh":
            layers.append(ns.Tanh())
    return ns.Sequential(*layers)


def _d_spec(slope, sn):
    c = "snconv" if sn else "conv"
    return [(c, 3, 128, 4, 2, 1), ("lrelu", slope), (c, 128, 256, 4, 2, 1), ("lrelu", slope), (c, 256, 512, 4, 2, 1),
            ("lrelu", slope), (c, 512, 1024, 4, 2, 1), ("lrelu", slope), ("conv", 1024, 19, 4, 1, 0)]


def _g_spec(slope):
    act = ("lrelu", slope) if slope is not None else ("relu",)
    return [("convT", 218, 1024, 4, 1, 0), ("convT", 1024, 512, 4, 2, 1), ("bn", 512), act,
            ("convT", 512, 256, 4, 2, 1), ("bn", 256), act, ("convT", 256, 128, 4, 2, 1), ("bn", 128), act,
            ("convT", 128, 3, 4, 2, 1), ("tanh",)]


SMALL = [
    ("dsprites_trunk", [("conv", 32, 32, 4, 2, 1), ("lrelu", 0.9), ("conv", 32, 64, 4, 2, 1), ("lrelu", 0.9),
                        ("conv", 64, 64, 4, 2, 1), ("lrelu", 0.9)], (6, 32, 32, 32)),
    ("dsprites_G", [("convT", 64, 64, 4, 2, 1), ("bn", 64), ("lrelu", 0.9), ("convT", 64, 64, 4, 2, 1), ("bn", 64),
                    ("lrelu", 0.9), ("convT", 64, 1, 4, 2, 1)], (6, 64, 4, 4)),
]


def _run(spec, in_shape, cuda, seed=0):
    import eadgan_b200.nn as enn
    os.environ["EADGAN_PRECISION"] = "bf16"
    torch.manual_seed(seed)
    ours = _build(enn, spec).to(cuda)
    ref, emu = _build(torch.nn, spec).to(cuda), _build(torch.nn, spec).to(cuda)
    ref.load_state_dict(ours.state_dict())
    emu.load_state_dict(ours.state_dict())
    x = torch.randn(*in_shape, device=cuda)
    xo, xr, xe = (x.clone().requires_grad_() for _ in range(3))
    yo, yr, ye = ours(xo), ref(xr), bf16_emul.emulate(ours, emu, xe)
    go = torch.randn_like(yr)
    names = ["x"] + [n for n, _ in ours.named_parameters()]
    gso = torch.autograd.grad(yo, [xo] + list(ours.parameters()), go)
    gsr = torch.autograd.grad(yr, [xr] + list(ref.parameters()), go)
    gse = torch.autograd.grad(ye, [xe] + list(emu.parameters()), go)
    # conv biases directly in front of a BatchNorm have a mathematically zero gradient: scale = sibling weight
    zero = {f"{i}.bias": f"{i}.weight" for i, it in enumerate(spec[:-1])
            if it[0] in ("conv", "convT") and spec[i + 1][0] == "bn"}
    return names, (yo, yr, ye), (gso, gsr, gse), zero


def _grad_errs(names, ga, gb, zero):
    """-> ({name: max-normalised error}, {name: L2-relative error}) of gradients ga against gb."""
    ref = dict(zip(names, gb))
    mx, l2 = {}, {}
    for n, a, b in zip(names, ga, gb):
        if n in zero:
            scale = float(ref[zero[n]].abs().max())
            mx[n] = float((a - b).abs().max()) / scale
        else:
            mx[n] = rel_err(a, b)
            l2[n] = _l2(a, b)
    return mx, l2


# Whole-stack bounds.  One bf16 rounding is 2^-9 = 0.2 % (rms 0.11 %) of an element; a stack output has been
# through 5 layers of rounded weights and rounded activations, so against the UN-rounded fp32 modules the error
# is a ~0.5-1 % rms noise whose maximum over 1e5 elements is 4-5 sigma.  Hence: L2-relative <= 2e-2 (the
# north_star bound as an rms statement) and tensor-normalised max <= 6e-2 for whole stacks; the literal 2e-2 MAX
# bound is asserted per layer on identical inputs (tests/test_tc_gpu.py: measured 2e-3..4e-3).
STACK_L2, STACK_MAX = 2e-2, 6e-2


@pytest.mark.parametrize("which,B", [("D", 8), ("D", 5), ("Dsn", 8), ("G", 8), ("G", 6)])
def test_gate_insensitive_stack_tight(cuda, which, B):
    """(1) every fused path of the chain vs the bf16-rounding reference, outputs AND every gradient."""
    spec = {"D": _d_spec(0.9, False), "Dsn": _d_spec(0.9, True), "G": _g_spec(0.9)}[which]
    shape = (B, 3, 64, 64) if which.startswith("D") else (B, 218, 1, 1)
    names, (yo, yr, ye), (gso, gsr, gse), zero = _run(spec, shape, cuda)
    assert yo.dtype == torch.float32 and yo.shape == yr.shape
    assert _l2(yo, ye) <= STACK_L2 and rel_err(yo, ye) <= STACK_MAX, (_l2(yo, ye), rel_err(yo, ye))
    assert _l2(yo, yr) <= STACK_L2 and rel_err(yo, yr) <= STACK_MAX, (_l2(yo, yr), rel_err(yo, yr))
    mx, l2 = _grad_errs(names, gso, gse, zero)
    assert max(l2.values()) <= STACK_L2, l2
    assert max(mx.values()) <= STACK_MAX, mx


@pytest.mark.parametrize("name,spec,shape", SMALL)
def test_small_channel_stacks(cuda, name, spec, shape):
    """dSprites-sized layers (32 / 64 channels): mixed tcgen05 + SIMT stages in one chain."""
    names, (yo, yr, ye), (gso, gsr, gse), zero = _run(spec, shape, cuda)
    assert _l2(yo, ye) <= STACK_L2 and rel_err(yo, ye) <= STACK_MAX
    mx, l2 = _grad_errs(names, gso, gse, zero)
    assert max(l2.values()) <= STACK_L2, l2
    assert max(mx.values()) <= STACK_MAX, mx


@pytest.mark.parametrize("which,B", [("D", 16), ("Dsn", 16), ("G", 16)])
def test_reference_stacks_vs_fp32(cuda, which, B):
    """(2) the real nets (ReLU / LeakyReLU(0.1) gates) against the un-rounded fp32 torch modules."""
    spec = {"D": _d_spec(0.1, False), "Dsn": _d_spec(0.1, True), "G": _g_spec(None)}[which]
    shape = (B, 3, 64, 64) if which.startswith("D") else (B, 218, 1, 1)
    names, (yo, yr, ye), (gso, gsr, gse), zero = _run(spec, shape, cuda)
    assert _l2(yo, yr) <= STACK_L2 and rel_err(yo, yr) <= STACK_MAX, (_l2(yo, yr), rel_err(yo, yr))
    mx, _ = _grad_errs(names, gso, gsr, zero)
    for n, a, b in zip(names, gso, gsr):
        if n in zero:
            assert mx[n] <= 2e-2, (n, mx[n])
            continue
        assert _cos(a, b) >= 0.98, (n, _cos(a, b))
        assert _l2(a, b) <= 0.2, (n, _l2(a, b))


def test_generator_module_and_running_stats(cuda):
    os.environ["EADGAN_PRECISION"] = "bf16"
    from eadgan_b200.steps.celeba import Generator
    from oracle.torch_oracle import CelebAGenerator
    torch.manual_seed(0)
    ours, ref = Generator().to(cuda), CelebAGenerator().to(cuda)
    ref.load_state_dict(ours.state_dict())
    B = 16
    z, code = torch.randn(B, 200, device=cuda), torch.rand(B, 8, device=cuda) * 2 - 1
    lab = torch.zeros(B, 10, device=cuda); lab[:, 3] = 1
    yo, yr = ours(z, lab, code), ref(z, lab, code)
    assert yo.shape == yr.shape and rel_err(yo, yr) <= 5e-2
    for k in ref.state_dict():
        if "running" in k:
            assert rel_err(ours.state_dict()[k], ref.state_dict()[k]) <= 1e-2, k
        if "num_batches" in k:
            assert int(ours.state_dict()[k]) == 1


def test_chain_matches_fp32_path(cuda):
    """same modules, both precisions of OUR implementation."""
    from eadgan_b200.steps.celeba import Discriminator
    torch.manual_seed(2)
    D = Discriminator().to(cuda).eval()  # eval: no power iteration, so both runs see the same weights
    x = torch.rand(6, 3, 64, 64, device=cuda) * 2 - 1
    os.environ["EADGAN_PRECISION"] = "fp32"
    a = D(x)[1]
    os.environ["EADGAN_PRECISION"] = "bf16"
    b = D(x)[1]
    assert rel_err(b, a) <= 2e-2
