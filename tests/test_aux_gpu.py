"""-m gpu: the remaining script paths of SURVEY.md section 8f rank 4 -- the pre-training loop of the MNIST affine
approximator (MNIST/approximate_rpqmnxy.py:111-153) and the eval-mode generator forwards of the inference scripts
(celebA/gen_imgs.py:106-200, MNIST/generate_image.py:146-154: load_state_dict -> .eval() -> G(z, labels, code))."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_approximator_pretraining_step(cuda, prec):
    from eadgan_b200.steps.approximator import ApproximatorStep
    from oracle import torch_oracle as O
    os.environ["EADGAN_PRECISION"] = prec
    st = O.build_approximator(seed=4, device=cuda, dtype=torch.float64)
    ours = ApproximatorStep(seed=4, device=cuda)
    for (k, a), (_, b) in zip(ours.A.state_dict().items(), st["A"].state_dict().items()):
        assert torch.equal(a, b.float()), k                       # same seeded init, same state_dict layout
    rs = np.random.RandomState(4)
    for it in range(3):
        code = O.sample_approximator(rs, 128)
        ref = O.step_approximator(st, code.double())
        rec = []
        out = ours(code.to(cuda), record=rec)
        assert abs(float(out["affine_loss"]) - ref["loss"]) <= 2e-5 * max(1.0, abs(ref["loss"])) * (1 if it == 0 else 50)
        if it == 0:      # later iterations carry Adam's lr * sign(g) noise
            for a, b in zip(rec[0]["grads"], ref["grads"]):
                assert rel_err(a, b) <= 1e-4


def test_eval_mode_generators(cuda):
    """inference: random state_dict loaded into OUR generators, .eval(), no_grad forward on 100 samples (the
    n_classes ** 2 grid of gen_imgs.py:116-122) vs the stock modules with the same state.  Eval-mode BatchNorm uses
    the running statistics; nothing may change them."""
    from eadgan_b200.steps import celeba as C, mnist as M
    from oracle import torch_oracle as O
    torch.manual_seed(11)
    for prec, tol in (("fp32", 2e-5), ("bf16", 2e-2)):
        os.environ["EADGAN_PRECISION"] = prec
        # CelebA
        ref = O.CelebAGenerator().to(cuda)
        for m in ref.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.3)
                m.running_var.uniform_(0.5, 1.5)
        ours = C.Generator().to(cuda)
        ours.load_state_dict(ref.state_dict())
        ref.eval(), ours.eval()
        z = torch.zeros(100, 200, device=cuda)
        z[:, :3] = torch.randn(100, 3, device=cuda)
        lab = torch.eye(10, device=cuda).repeat(10, 1)
        code = torch.zeros(100, 8, device=cuda)
        code[:, 0] = torch.linspace(-1, 1, 100, device=cuda)
        before = {k: v.clone() for k, v in ours.state_dict().items()}
        with torch.no_grad():
            a, b = ours(z, lab, code), ref(z, lab, code)
        assert rel_err(a, b) <= tol, (prec, rel_err(a, b))
        assert all(torch.equal(v, before[k]) for k, v in ours.state_dict().items())
        # MNIST
        refm = O.MnistGenerator().to(cuda)
        refm.apply(O.mnist_weights_init_normal)
        oursm = M.Generator().to(cuda)
        oursm.load_state_dict(refm.state_dict())
        refm.eval(), oursm.eval()
        zm = torch.randn(100, 62, device=cuda)
        cm = torch.rand(100, 7, device=cuda) * 2 - 1
        with torch.no_grad():
            assert rel_err(oursm(zm, lab, cm), refm(zm, lab, cm)) <= tol
