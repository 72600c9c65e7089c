"""-m gpu: BASELINE configs[2] -- the colored-dSprites stage-2 step (colored_dSprites/rp_color.py) through the
drop-in modules vs the oracle restatement, which is pinned to the reference script itself
(tests/golden/colored_*.json, tests/test_cpu.py)."""
import pytest
import torch

import step_util as U
from conftest import rel_err

pytestmark = pytest.mark.gpu
LOSSES = ("d_loss", "g_loss", "cat_loss", "cont_loss", "affine_loss", "relative_cat_loss", "total")


def test_state_dict_layout_matches_oracle(cuda):
    from eadgan_b200.steps.colored import ColoredDSpritesStep
    from oracle import torch_oracle as O
    ours = ColoredDSpritesStep(seed=3, device=cuda, pxy_state=O.dsprites_pxy_state(3, colored=True))
    st = O.build_dsprites(seed=3, device=cuda, colored=True)
    for key, net in (("G", ours.G), ("D", ours.D), ("E", ours.E), ("Epxy", ours.Epxy)):
        a, b = net.state_dict(), st[key].state_dict()
        assert list(a.keys()) == list(b.keys()), key
        for k in a:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, (key, k)
            assert torch.equal(a[k], b[k]), (key, k)
    assert ours.opt_info.param_groups[0]["lr"] == st["opt_info"].param_groups[0]["lr"] == 0.0002


def test_colorize_matches_float64_product(cuda):
    from eadgan_b200.steps.colored import colorize
    from oracle import torch_oracle as O
    import numpy as np
    img = O.synth_dsprites_images(8, 1).to(cuda)
    gains = O.sample_colored(np.random.RandomState(1), 8)["color"].to(cuda)
    want = (img.unsqueeze(1).repeat(1, 3, 1, 1) * gains).float()
    assert torch.equal(colorize(img, gains), want)


def test_colored_step_fp32(cuda):
    ref, rec, losses, st, ours = U.run_pair_colored(cuda, 16, "fp32")
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
def test_colored_step_bf16(cuda, B):
    """bf16 tcgen05 chain (configs[2] runs at 512 global = 256 per GPU on two GPUs)."""
    ref, rec, losses, st, ours = U.run_pair_colored(cuda, B, "bf16")
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
