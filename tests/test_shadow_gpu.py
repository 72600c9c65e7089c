"""-m gpu: the shadow ``utils_*.py`` modules (eadgan_b200/shadow) against the oracle's restatements of the reference's
helper functions (oracle/torch_oracle.py, pinned to celebA/utils_rpqxy.py, dSprites/utils_rp.py, dSprites/utils_pxy.py,
colored_dSprites/utils_rp_color.py, MNIST/utils_rpqmnxy.py by tests/test_cpu.py): values in fp64 on the CPU vs the
device-side fp32 evaluation, and gradients through the regularisers."""
import importlib.util
import os

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(sub, name):
    spec = importlib.util.spec_from_file_location(f"shadow_{sub}_{name}", os.path.join(ROOT, "eadgan_b200", "shadow", sub, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _codes(n, k, seed, cuda):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n, k, generator=g) * 2 - 1
    return c.double(), c.to(cuda)


@pytest.mark.parametrize("case", ["celebA", "dSprites", "colored"])
def test_matrix_and_regulariser(cuda, case):
    from oracle import torch_oracle as O
    if case == "celebA":
        m, k = _load("celebA", "utils_rpqxy"), 5
        mat, ref_mat, reg, ref_reg = m.get_matrix, O.celeba_get_matrix, m.affine_regularzier, O.celeba_affine_regularizer
    elif case == "dSprites":
        m, k = _load("dSprites", "utils_rp"), 4
        mat, ref_mat, reg, ref_reg = m.get_matrix_D, O.dsprites_get_matrix, m.affine_regularzier, O.dsprites_affine_regularizer
    else:
        m, k = _load("colored_dSprites", "utils_rp_color"), 7
        mat, ref_mat = m.get_matrix, O.dsprites_get_matrix
        reg, ref_reg = m.affine_color_regularzier, O.colored_affine_color_regularizer
    a64, a32 = _codes(33, k, 1, cuda)
    b64, b32 = _codes(33, k, 2, cuda)
    ka = 4 if case == "colored" else k
    assert rel_err(mat(a32[:, :ka]), ref_mat(a64[:, :ka])) <= 2e-6
    a32.requires_grad_(), b32.requires_grad_(), a64.requires_grad_(), b64.requires_grad_()
    out, ref = reg(a32, b32), ref_reg(a64, b64)
    assert out.dtype == torch.float32 and rel_err(out, ref) <= 2e-5
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(3))
    (out * w.to(cuda)).sum().backward()
    (ref * w.double()).sum().backward()
    assert rel_err(a32.grad, a64.grad) <= 2e-4 and rel_err(b32.grad, b64.grad) <= 2e-4
    # parameter <-> code maps are inverse to each other and leave further columns zero
    fwd = m.from_latent_vector_2_affine_para(a32.detach()[:, :ka]) if hasattr(m, "from_latent_vector_2_affine_para") else None
    if fwd is not None:
        assert rel_err(m.from_affine_para_2_latent_vector(fwd), a32.detach()[:, :ka]) <= 1e-5


def test_stage1_helpers(cuda):
    from oracle import torch_oracle as O
    for sub, k in (("dSprites", 3), ("colored_dSprites", 6)):
        m = _load(sub, "utils_pxy")
        a64, a32 = _codes(17, k, 4, cuda)
        b64, b32 = _codes(17, k, 5, cuda)
        assert rel_err(m.get_matrix_pxy(a32[:, :3]), O.pxy_get_matrix(a64[:, :3])) <= 2e-6
        assert rel_err(m.get_matrix_pxy_align(a32[:, :3]), O.dsprites_align_matrix(a64[:, :3])) <= 2e-6
        assert rel_err(m.affine_regularzier_pxy(a32, b32), O.pxy_affine_regularizer(a64, b64)) <= 2e-5
        e = m.get_enlarge_matrix(a32[:, :3])
        assert tuple(e.shape) == (17, 3, 3) and float(e[0, 0, 0]) == pytest.approx(0.6) and float(e[0, 2, 2]) == 1.0


def test_closed_form_inverse_patch(cuda):
    """patch() rebinds torch.inverse: CUDA [B,3,3] through the adjugate (no host sync), everything else stock."""
    import importlib
    P = importlib.import_module("eadgan_b200.patch")
    m = torch.randn(9, 3, 3, device=cuda, dtype=torch.float64) + 2 * torch.eye(3, device=cuda, dtype=torch.float64)
    want = torch.inverse(m)
    P.patch()
    try:
        assert torch.inverse is P.inverse
        assert rel_err(torch.inverse(m), want) <= 1e-10
        big = torch.randn(4, 4, device=cuda, dtype=torch.float64)
        assert rel_err(torch.inverse(big), torch.linalg.inv(big)) <= 1e-10
    finally:
        P.unpatch()
    assert torch.inverse is not P.inverse
