"""-m gpu: BASELINE configs[1] -- the dSprites stage-2 step (dSprites/rp.py: frozen Encoder_pxy alignment,
D / G / Encoder, two phases, two Adams) through the drop-in modules vs the oracle restatement, which is pinned
bit-for-bit to the reference script itself (tests/golden/dsprites_*.json, tests/test_cpu.py)."""
import pytest
import torch

import step_util as U
from conftest import rel_err

pytestmark = pytest.mark.gpu
LOSSES = ("d_loss", "g_loss", "cat_loss", "cont_loss", "affine_loss", "relative_cat_loss", "total")


def test_state_dict_layout_matches_oracle(cuda):
    from eadgan_b200.steps.dsprites import DSpritesStep
    from oracle import torch_oracle as O
    ours = DSpritesStep(seed=3, device=cuda, pxy_state=O.dsprites_pxy_state(3))
    st = O.build_dsprites(seed=3, device=cuda)
    for key, net in (("G", ours.G), ("D", ours.D), ("E", ours.E), ("Epxy", ours.Epxy)):
        a, b = net.state_dict(), st[key].state_dict()
        assert list(a.keys()) == list(b.keys()), key
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, (key, k)
            assert torch.equal(a[k], b[k]), (key, k)       # same seeded construction order -> same init
        st[key].load_state_dict(a)
        net.load_state_dict(b)


def test_dsprites_step_fp32(cuda):
    ref, rec, losses, st, ours = U.run_pair_dsprites(cuda, 16, "fp32")
    for k in LOSSES:
        assert abs(losses[k] - ref["losses"][k]) <= 2e-5 * max(1.0, abs(ref["losses"][k])), (k, losses, ref["losses"])
    names = U.dsprites_grad_names(ours)
    for ph in range(2):
        errs = U.phase_errors(names[ph], rec[ph]["grads"], ref["phases"][ph]["grads"], U.DSPRITES_ZERO_GRAD)
        mx = sorted(v[0] for v in errs.values())
        assert mx[-1] <= 1e-2, (ph, errs)
        assert mx[len(mx) // 2] <= 5e-3, (ph, errs)
    so, sr = ours.G.state_dict(), st["G"].state_dict()
    for k in sr:
        if "running" in k:
            assert rel_err(so[k], sr[k]) <= 1e-4, k


@pytest.mark.parametrize("B", [32, 256])
def test_dsprites_step_bf16(cuda, B):
    """bf16 tcgen05 chain on the 32/64-channel trunks (configs[1] runs at batch 256)."""
    ref, rec, losses, st, ours = U.run_pair_dsprites(cuda, B, "bf16")
    for k in LOSSES:
        assert abs(losses[k] - ref["losses"][k]) <= 2e-2 * max(1.0, abs(ref["losses"][k])), (k, losses, ref["losses"])
    names = U.dsprites_grad_names(ours)
    for ph in range(2):
        errs = U.phase_errors(names[ph], rec[ph]["grads"], ref["phases"][ph]["grads"], U.DSPRITES_ZERO_GRAD)
        for n, (mx, l2, cs) in errs.items():
            if cs is None:
                assert mx <= 2e-2, (ph, n, mx)
            elif cs == "small":
                assert mx <= 0.15, (ph, n, mx)
            else:
                assert cs >= 0.95, (ph, n, cs)
                assert l2 <= 0.35, (ph, n, l2)
