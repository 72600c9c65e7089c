"""Test helper: forced activation gates (VERDICT r01 "what's weak" 2).

A ReLU / LeakyReLU network is piecewise linear in its parameters' gradients: a single gate that differs between
two evaluations (a pre-activation within rounding of 0) moves EVERY upstream gradient by a finite amount, whatever
the quality of the kernels.  To hold every gradient tensor of the bf16 tensor-core step to north_star's 2e-2, the
oracle is therefore evaluated on the SAME piecewise-linear branch as our run:

  1. our step runs with ``eadgan_b200.chain.gate_log`` switched on: every conv stack records, per ReLU / LeakyReLU
     stage, the mask ``saved output > 0`` -- the very bits its fused backward kernels read;
  2. the oracle's ``nn.ReLU`` / ``nn.LeakyReLU`` modules are replaced by ``ForcedGate`` modules that consume those
     masks in call order: y = where(mask, x, slope * x).  Where the oracle's own gate agrees (all but ~0.1 % of the
     elements) this IS ReLU / LeakyReLU; where it does not, |x| is within bf16 rounding of 0, so the forward value
     moves by less than one bf16 ulp of the activation scale.

What is left between the two runs is rounding of smooth arithmetic, which the 2e-2 bound is about."""
import numpy as np
import torch
import torch.nn as tnn


class ForcedGate(tnn.Module):
    def __init__(self, slope, queue):
        super().__init__()
        self.slope, self.queue = float(slope), queue
        self.flips = 0
        self.total = 0

    def forward(self, x):
        mask = self.queue.pop(0)
        self.flips += int(((x > 0) != mask).sum())
        self.total += mask.numel()
        return torch.where(mask, x, x * self.slope)


def install(seq, queue):
    """replace every ReLU / LeakyReLU of the torch Sequential ``seq`` by a ForcedGate fed from ``queue``"""
    gates = []
    for name, m in list(seq._modules.items()):
        if isinstance(m, tnn.LeakyReLU):
            seq._modules[name] = ForcedGate(m.negative_slope, queue)
        elif isinstance(m, tnn.ReLU):
            seq._modules[name] = ForcedGate(0.0, queue)
        else:
            continue
        gates.append(seq._modules[name])
    return gates


def celeba_forced(dev, B, precision="bf16", seed=0, oracle_dtypes=(torch.float64,)):
    """-> dict(forced=[oracle record per dtype in ``oracle_dtypes``], ours=our per-phase record, losses=our losses,
    free=free oracle record (first dtype), flips=..., gates=..., step=our step object).

    Runs on identical seeded inputs / weights: the free oracle (its post-phase states are the common starting point
    of every later phase of every run), OUR step (gates logged), then the oracle with OUR gates forced, once per
    requested dtype (fp64 = the referee; fp32 = stock torch's own rounding error on the same branch)."""
    import os
    from eadgan_b200 import chain
    from eadgan_b200.steps.celeba import CelebAStep
    from oracle import torch_oracle as O
    os.environ["EADGAN_PRECISION"] = precision
    imgs = O.synth_celeba_images(B, seed).to(dev)
    draws = O.sample_celeba(np.random.RandomState(seed), B)
    dt0 = oracle_dtypes[0]

    free = O.step_celeba(O.build_celeba(seed=seed, device=dev, dtype=dt0), imgs.to(dt0), draws)
    states = [free["phases"][i]["state_after"] for i in range(2)]

    def cast(sd, dt):
        return {k: v.to(dt) if v.is_floating_point() else v for k, v in sd.items()}

    ours = CelebAStep(seed=seed, device=dev)

    def load_ours(i):
        ours.G.load_state_dict(cast(states[i]["G"], torch.float32))
        ours.D.load_state_dict(cast(states[i]["D"], torch.float32))

    rec = []
    chain.gate_log = []
    try:
        losses = ours(imgs, draws["z"].to(dev), draws["code"].to(dev), draws["labels"].to(dev), record=rec,
                      after_phase=load_ours)
        log = chain.gate_log
    finally:
        chain.gate_log = None
    qG = [m for seq, masks in log if seq is ours.G.conv_blocks for m in masks]
    qD = [m for seq, masks in log if seq is ours.D.main for m in masks]
    assert len(qG) == 2 * 3 and len(qD) == 6 * 4, (len(qG), len(qD))

    forced, flips, total = [], 0, 0
    for dt in oracle_dtypes:
        st = O.build_celeba(seed=seed, device=dev, dtype=dt)
        g_q, d_q = list(qG), list(qD)
        gates = install(st["G"].conv_blocks, g_q) + install(st["D"].main, d_q)

        def load_ref(i, _st=st, _dt=dt):
            _st["G"].load_state_dict(cast(states[i]["G"], _dt))
            _st["D"].load_state_dict(cast(states[i]["D"], _dt))

        forced.append(O.step_celeba(st, imgs.to(dt), draws, after_phase=load_ref))
        assert not g_q and not d_q
        if dt == dt0:
            flips, total = sum(g.flips for g in gates), sum(g.total for g in gates)
        del st
    return {"forced": forced, "ours": rec, "losses": {k: float(v) for k, v in losses.items()}, "free": free,
            "flips": flips, "gates": total, "step": ours}
