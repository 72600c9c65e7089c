"""-m gpu: the bf16 tcgen05 chain executor (eadgan_b200.chain) on whole G / D conv stacks against the
stock torch fp32 modules with identical weights.  Tolerance: north_star bf16 bound 2e-2 on
outputs; gradients (which pass through several bf16 layers) 4e-2, tensor-normalised."""
import os

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _err(a, b):
    return rel_err(a, b) if float(b.abs().max()) > 1e-6 else float((a - b).abs().max())


@pytest.mark.parametrize("B", [4, 7, 32])
def test_generator_chain(cuda, B):
    os.environ["EADGAN_PRECISION"] = "bf16"
    from eadgan_b200.steps.celeba import Generator
    from oracle.torch_oracle import CelebAGenerator
    torch.manual_seed(0)
    ours, ref = Generator().to(cuda), CelebAGenerator().to(cuda)
    ref.load_state_dict(ours.state_dict())
    z = torch.randn(B, 200, device=cuda)
    lab = torch.zeros(B, 10, device=cuda); lab[:, 3] = 1
    code = torch.rand(B, 8, device=cuda) * 2 - 1
    yo, yr = ours(z, lab, code), ref(z, lab, code)
    assert yo.shape == yr.shape and yo.dtype == torch.float32
    assert rel_err(yo, yr) <= 2e-2
    go = torch.randn_like(yr)
    po, pr = list(ours.parameters()), list(ref.parameters())
    gso = torch.autograd.grad(yo, po, go)
    gsr = torch.autograd.grad(yr, pr, go)
    errs = [_err(a, b) for a, b in zip(gso, gsr)]
    assert max(errs) <= 4e-2, errs
    for k in ref.state_dict():
        if "running" in k:
            assert rel_err(ours.state_dict()[k], ref.state_dict()[k]) <= 1e-2, k


@pytest.mark.parametrize("B", [4, 9, 32])
def test_discriminator_chain(cuda, B):
    os.environ["EADGAN_PRECISION"] = "bf16"
    from eadgan_b200.steps.celeba import Discriminator
    from oracle.torch_oracle import CelebADiscriminator
    torch.manual_seed(1)
    ours, ref = Discriminator().to(cuda), CelebADiscriminator().to(cuda)
    ref.load_state_dict(ours.state_dict())
    x = (torch.rand(B, 3, 64, 64, device=cuda) * 2 - 1)
    xo, xr = x.clone().requires_grad_(), x.clone().requires_grad_()
    (co, to_, vo), (cr, tr, vr) = ours(xo), ref(xr)
    assert rel_err(vo, vr) <= 2e-2 and rel_err(to_, tr) <= 2e-2 and rel_err(co, cr) <= 2e-2
    lo = (vo.sum() + (to_ ** 2).sum() + co[:, 0].sum())
    lr = (vr.sum() + (tr ** 2).sum() + cr[:, 0].sum())
    po = [xo] + list(ours.parameters())
    pr = [xr] + list(ref.parameters())
    gso, gsr = torch.autograd.grad(lo, po), torch.autograd.grad(lr, pr)
    errs = [_err(a, b) for a, b in zip(gso, gsr)]
    assert max(errs) <= 4e-2, errs
    for k in ("main.0.weight_u", "main.6.weight_v"):
        assert rel_err(ours.state_dict()[k], ref.state_dict()[k]) <= 1e-4


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
