"""-m gpu: the whole CelebA step captured in a CUDA graph replays to the same results as eager execution."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_graphed_step_matches_eager(cuda, prec):
    os.environ["EADGAN_PRECISION"] = prec
    from eadgan_b200.graph import GraphedStep
    from eadgan_b200.steps.celeba import CelebAStep
    from oracle.torch_oracle import sample_celeba, synth_celeba_images
    B, W, K = 16, 2, 3

    def batch(i):
        d = sample_celeba(np.random.RandomState(10 + i), B)
        return [synth_celeba_images(B, i).to(cuda), d["z"].to(cuda), d["code"].to(cuda), d["labels"].to(cuda)]

    eager = CelebAStep(seed=0, device=cuda)
    graphed = CelebAStep(seed=0, device=cuda)
    gs = GraphedStep(graphed, batch(0), warmup=W)       # W eager warm-up steps on batch(0) inside
    for _ in range(W):
        eager(*batch(0))
    for i in range(1, K + 1):
        le = {k: float(v) for k, v in eager(*batch(i)).items()}
        lg = {k: float(v) for k, v in gs(*batch(i)).items()}
        # both paths are run-to-run deterministic up to the fp64 atomics of the BatchNorm sums (1e-16 relative):
        # the tcgen05 kernels use no float atomics, and the SIMT wgrad sums its split partials in a fixed order
        tol = 1e-6 if prec == "bf16" else 1e-4
        for k in le:
            assert abs(le[k] - lg[k]) <= tol * max(1.0, abs(le[k])), (i, k, le, lg)
    # weights after W + K optimiser steps: identical up to Adam's lr*sign(g) noise on elements whose gradient
    # is at the level of the atomics' summation-order noise (at most a few steps of lr = 2e-4 / 1e-3)
    for net_e, net_g in ((eager.D, graphed.D), (eager.G, graphed.G)):
        for (n, a), (_, b) in zip(net_e.state_dict().items(), net_g.state_dict().items()):
            if a.is_floating_point() and not (prec == "fp32" and n.endswith(("_u", "_v"))):
                d = (a - b).abs()
                assert float(d.max()) <= 6e-3, (n, float(d.max()))
                assert float(d.mean()) <= (2e-5 if prec == "bf16" else 3e-4), (n, float(d.mean()))
    # python-side step mirrors follow the device counter
    st = next(iter(graphed.opt_G.state.values()))
    assert st["step"] == W + K
    assert int(graphed.opt_G._step_dev) == W + K
    assert gs.kernels_per_replay > 100


def test_prefetch_pipeline_gives_the_same_steps(cuda):
    """GraphedStep(..., prefetch=next batch): the next step's host->device copy runs on a copy stream during the
    current replay; results must equal feeding the same batches without the pipeline."""
    os.environ["EADGAN_PRECISION"] = "bf16"
    from eadgan_b200.graph import GraphedStep
    from eadgan_b200.steps.celeba import CelebAStep
    from oracle.torch_oracle import sample_celeba, synth_celeba_images
    B = 16

    def host_batch(i):
        d = sample_celeba(np.random.RandomState(20 + i), B)
        return [t.contiguous().pin_memory() for t in (synth_celeba_images(B, i), d["z"], d["code"], d["labels"])]

    batches = [host_batch(i) for i in range(4)]
    dev0 = [t.to(cuda) for t in batches[0]]
    plain = GraphedStep(CelebAStep(seed=0, device=cuda), dev0, warmup=1)
    piped = GraphedStep(CelebAStep(seed=0, device=cuda), dev0, warmup=1)
    for i in range(4):
        a = {k: float(v) for k, v in plain(*batches[i]).items()}
        b = {k: float(v) for k, v in piped(*batches[i], prefetch=batches[(i + 1) % 4]).items()}
        assert a == b, (i, a, b)
