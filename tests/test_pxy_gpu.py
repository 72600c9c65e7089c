"""-m gpu: stage-1 Encoder_pxy pre-training step (dSprites/pxy.py, colored_dSprites/pxy_color.py) vs the oracle."""
import os

import numpy as np
import pytest
import torch

import step_util as U

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("colored", [False, True])
@pytest.mark.parametrize("prec,B", [("fp32", 16), ("bf16", 128)])
def test_pxy_step(cuda, colored, prec, B):
    from eadgan_b200.steps.pxy import PxyStep
    from oracle import torch_oracle as O
    os.environ["EADGAN_PRECISION"] = prec
    imgs = O.synth_dsprites_images(B, 0).to(cuda)
    draws = O.sample_pxy(np.random.RandomState(0), B, colored)
    st = O.build_pxy(seed=0, device=cuda, dtype=torch.float64, colored=colored)
    ref = O.step_pxy(st, imgs, draws)
    ours = PxyStep(seed=0, device=cuda, colored=colored)
    rec = []
    out = ours(imgs, draws["code"].to(cuda), draws["color"].to(cuda) if colored else None, record=rec)
    tol = 2e-5 if prec == "fp32" else 2e-2
    want = ref["losses"]["affine_loss"]
    assert abs(float(out["affine_loss"]) - want) <= tol * max(1.0, abs(want))
    names = ["E." + n for n, _ in ours.E.named_parameters()]
    errs = U.phase_errors(names, rec[0]["grads"], ref["phases"][0]["grads"], {})
    for n, (mx, l2, cs) in errs.items():
        if prec == "fp32":
            # one LeakyReLU(0.1) gate flip (pre-activation within fp32 rounding of 0) moves a bias gradient by
            # ~1e-2 of its scale: bound the worst tensor loosely, the median tightly
            assert mx <= 3e-2, (n, mx)
        elif cs in (None, "small"):
            assert mx <= 0.15, (n, mx)
        else:
            assert cs >= 0.95 and l2 <= 0.35, (n, cs, l2)
    if prec == "fp32":
        mxs = sorted(v[0] for v in errs.values())
        assert mxs[len(mxs) // 2] <= 5e-3, errs
