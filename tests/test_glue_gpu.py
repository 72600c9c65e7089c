"""-m gpu: the fused affine-glue kernels (csrc/glue.cu) against the differentiable torch restatements in
eadgan_b200/affine.py (which tests/test_cpu.py pins to the reference's utils_*.py through the oracle) and against
the stock F.affine_grid + F.grid_sample pair."""
import pytest
import torch
import torch.nn.functional as TF

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["celeba", "dsprites", "mnist"])
def test_relative_code_value_and_gradient(cuda, mode):
    from eadgan_b200 import affine
    fused, ref, k = {"celeba": (affine.celeba_relative_code, affine.celeba_relative_code_torch, 5),
                     "dsprites": (affine.dsprites_relative_code, affine.dsprites_relative_code_torch, 4),
                     "mnist": (affine.mnist_relative_rows, affine.mnist_relative_rows_torch, 7)}[mode]
    g = torch.Generator(device="cpu").manual_seed(11)
    wide = torch.rand(257, 19, generator=g) * 2 - 1          # codes arrive as column slices of a wider head output
    r1 = wide[:, 1:1 + k].to(cuda).double().requires_grad_(True)
    t1 = (torch.rand(257, k, generator=g) * 2 - 1).to(cuda).double().requires_grad_(True)
    want = ref(r1, t1)                                        # fp64 referee
    go = torch.randn(want.shape, generator=g).to(cuda)
    want.backward(go.double())
    base = wide.to(cuda).requires_grad_(True)
    r2 = base[:, 1:1 + k]
    t2 = t1.detach().float().requires_grad_(True)
    got = fused(r2, t2)
    assert got.dtype == torch.float32 and got.shape == want.shape
    got.backward(go)
    assert (got.double() - want).abs().max() <= 2e-5 * max(1.0, float(want.abs().max()))
    gr = base.grad[:, 1:1 + k].double()
    assert float(base.grad[:, 0].abs().max()) == 0 and float(base.grad[:, 1 + k:].abs().max()) == 0
    for a, b in ((gr, r1.grad), (t2.grad.double(), t1.grad)):
        assert (a - b).abs().max() <= 2e-4 * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("shape", [(7, 3, 64, 64), (5, 1, 64, 64), (4, 1, 32, 32)])
@pytest.mark.parametrize("padding", ["border", "zeros"])
def test_stn_matches_affine_grid_plus_grid_sample(cuda, shape, padding):
    from eadgan_b200 import affine
    torch.manual_seed(3)
    img = torch.rand(shape, device=cuda) * 2 - 1
    n = shape[0]
    code = torch.rand(n, 5, device=cuda) * 2 - 1
    theta = affine.celeba_matrix(code)[:, 0:2].contiguous()
    theta[0] = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], device=cuda)        # identity
    theta[1] = torch.tensor([[1.3, 0.2, 0.9], [-0.1, 0.8, -0.7]], device=cuda)      # samples far outside
    grid = TF.affine_grid(theta, list(img.shape), align_corners=False)
    want = TF.grid_sample(img, grid, padding_mode=padding, align_corners=False)
    got = affine.stn(img, theta, padding_mode=padding)
    assert (got - want).abs().max() <= 5e-5
    assert (got[0] - img[0]).abs().max() <= 1e-5
    # with a gradient requested the stock differentiable ops are used
    th = theta.clone().requires_grad_(True)
    out = affine.stn(img, th, padding_mode=padding)
    assert out.requires_grad


@pytest.mark.parametrize("padding", ["border", "zeros"])
@pytest.mark.parametrize("shape", [(5, 3, 64, 64), (4, 1, 32, 48)])
def test_affine_grid_and_grid_sample_forward_and_backward(cuda, padding, shape):
    """F.affine_grid / F.grid_sample as separate operators (what the unmodified scripts call) against stock torch:
    values, d/d(input), d/d(grid) through to d/d(theta).  The transforms push part of the grid outside [-1, 1] so
    that the 'border' clamp (zero coordinate gradient) and the 'zeros' corners are exercised.  dSprites/rp.py:200-211,
    374-377,399-400; colored_dSprites/pxy_color.py:82-96."""
    import torch.nn.functional as TF
    from eadgan_b200 import functional as Fn
    n, c, h, w = shape
    torch.manual_seed(21)
    img = torch.randn(n, c, h, w, device=cuda)
    theta = torch.eye(2, 3, device=cuda).repeat(n, 1, 1) + 0.35 * torch.randn(n, 2, 3, device=cuda)
    gout = torch.randn(n, c, h, w, device=cuda)
    a_img, a_th = img.clone().requires_grad_(), theta.clone().requires_grad_()
    b_img, b_th = img.double().requires_grad_(), theta.double().requires_grad_()
    grid = Fn.affine_grid(a_th, (n, c, h, w))
    ours = Fn.grid_sample(a_img, grid, padding_mode=padding)
    ref_grid = TF.affine_grid(b_th, [n, c, h, w], align_corners=False)
    ref = TF.grid_sample(b_img, ref_grid, padding_mode=padding, align_corners=False)
    assert rel_err(grid, ref_grid) <= 1e-6
    assert rel_err(ours, ref) <= 2e-5
    ours.backward(gout)
    ref.backward(gout.double())
    assert rel_err(a_img.grad, b_img.grad) <= 2e-5
    # the coordinate gradient is discontinuous where a sample point crosses a pixel boundary: fp32 vs fp64 place a
    # handful of points on different sides, which moves d theta by a few 1e-4 of its size; bounded at 2e-3
    assert rel_err(a_th.grad, b_th.grad) <= 2e-3
    # the fused forward-only kernel and the two-operator path agree bit for bit
    from eadgan_b200 import affine
    assert torch.equal(affine.stn(img, theta, padding_mode=padding), ours.detach())
