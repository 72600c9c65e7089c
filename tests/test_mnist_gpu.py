"""-m gpu: BASELINE configs[0] -- the MNIST step (MNIST/EAD-GAN_rpqmnxy.py: G with Upsample + 3x3 convs +
BatchNorm eps 0.8, spectral-normalised D / Encoder, LSGAN losses, three phases, three Adams) through the drop-in
modules vs the oracle restatement, which is pinned to the reference script (tests/golden/mnist_*.json)."""
import pytest
import torch

import step_util as U
from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_state_dict_layout_and_init_match_oracle(cuda):
    """same seeded construction order + weights_init_normal (which reaches the spectral-normalised convs through
    the weight / weight_orig storage aliasing) -> bit-identical initial state."""
    from eadgan_b200.steps.mnist import MnistStep
    from oracle import torch_oracle as O
    ours = MnistStep(seed=2, device=cuda, approximator_state=O.mnist_approximator_state(2))
    st = O.build_mnist(seed=2, device=cuda)
    for key, net in (("G", ours.G), ("D", ours.D), ("E", ours.E), ("A", ours.A)):
        a, b = net.state_dict(), st[key].state_dict()
        assert list(a.keys()) == list(b.keys()), key
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, (key, k)
            assert torch.equal(a[k], b[k]), (key, k)
        st[key].load_state_dict(a)
        net.load_state_dict(b)
    assert ours.G.conv_blocks[3].eps == 0.8 and ours.G.conv_blocks[0].eps == 1e-5


@pytest.mark.parametrize("prec,B", [("fp32", 16), ("fp32", 64), ("bf16", 64)])
def test_mnist_step(cuda, prec, B):
    """no layer of this configuration is k4 s2 p1, so every conv runs on the any-geometry SIMT kernels; in bf16
    mode the D trunk (a pure Conv + LeakyReLU Sequential) still goes through the chain executor, i.e. its
    intermediate activations are stored as bf16 -> north_star's 2e-2 bound there, 1e-5-class bounds in fp32."""
    ref, rec, losses, st, ours = U.run_pair_mnist(cuda, B, prec)
    ltol = 2e-5 if prec == "fp32" else 2e-2
    for k in ("g_loss", "d_loss", "info_loss"):
        assert abs(losses[k] - ref["losses"][k]) <= ltol * max(1.0, abs(ref["losses"][k])), (k, losses, ref["losses"])
    names = U.mnist_grad_names(ours)
    for ph in range(3):
        ours_g = [g for g in rec[ph]["grads"]]
        ref_g = ref["phases"][ph]["grads"]
        keep = [i for i, g in enumerate(ref_g) if g is not None]
        assert all(ours_g[i] is not None for i in keep), ph
        # parameters that never receive a gradient (the Encoder's noise head) must not get one from us either
        assert all(ours_g[i] is None for i, g in enumerate(ref_g) if g is None), ph
        errs = U.phase_errors([names[ph][i] for i in keep], [ours_g[i] for i in keep], [ref_g[i] for i in keep],
                              U.MNIST_ZERO_GRAD)
        if prec == "fp32":
            mx = sorted(v[0] for v in errs.values())
            assert mx[-1] <= 1e-2, (ph, errs)
            assert mx[len(mx) // 2] <= 5e-3, (ph, errs)
        else:
            for n, (mx, l2, cs) in errs.items():
                if cs in (None, "small"):
                    assert mx <= 0.15, (ph, n, mx)
                else:
                    assert cs >= 0.95 and l2 <= 0.35, (ph, n, cs, l2)
    for key, net in (("G", ours.G), ("E", ours.E)):
        so, sr = net.state_dict(), st[key].state_dict()
        for k in sr:
            if "running" in k:
                assert rel_err(so[k], sr[k]) <= (1e-4 if prec == "fp32" else 2e-2), (key, k)
            if "num_batches" in k:
                assert int(so[k]) == int(sr[k]), (key, k)
