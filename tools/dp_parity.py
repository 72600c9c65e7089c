"""N-rank data-parallel step  vs  single-device step on the same global batch.  Used by
bench.py (``dp_parity`` block when WORLD_SIZE > 1) and tools/dp_check.py (torchrun script / tests/test_dp_gpu.py)."""
import json
import os

import numpy as np
import torch
import torch.distributed as dist

from eadgan_b200 import functional as Fn, parallel
from eadgan_b200.steps.celeba import CelebAStep
# conv biases in front of a train-mode BatchNorm: mathematically zero gradient (what a run holds there is rounding
# noise), compared on the scale of the sibling weight gradient -- the same rule as tests/step_util.py
_ZERO_GRAD = {"G.conv_blocks.1.bias": "G.conv_blocks.1.weight", "G.conv_blocks.4.bias": "G.conv_blocks.4.weight",
              "G.conv_blocks.7.bias": "G.conv_blocks.7.weight"}


def dp_parity(B_local, dev, seed=0):
    """N-rank data-parallel step on the global batch  vs  single-device step on the same batch, both OUR
    implementation in the current EADGAN_PRECISION (SURVEY.md section 8e "parity definition").  Called by every rank
    (bench.py when WORLD_SIZE > 1, and this script); returns the error report on rank 0, None elsewhere.

    Rank 0 first runs the single-device step (data parallelism detached) and keeps its per-phase gradients, losses and
    post-phase states; the states are broadcast, and every phase of the N-rank run starts from them (so Adam's
    lr * sign(g) noise of one phase does not leak into the next: all three phases are compared from common weights).
    Compared: the three losses (mean over ranks), every all-reduced + averaged gradient of every phase, BatchNorm
    running statistics after the step."""
    from eadgan_b200 import chain, synthetic
    dp = parallel.get()
    # the phases restart from state snapshots taken between them: run every spectral-norm power iteration in line, so
    # that a snapshot holds u / v exactly as far as the forwards got (a prefetch runs up to three iterations ahead)
    saved_prefetch, chain.prefetch_enabled = chain.prefetch_enabled, False
    try:
        return _dp_parity(B_local, dev, seed, dp, synthetic)
    finally:
        chain.prefetch_enabled = saved_prefetch


def _dp_parity(B_local, dev, seed, dp, synthetic):
    world, rank = dp.world_size, dp.rank
    Bg = B_local * world
    imgs = synthetic.celeba_images(Bg, seed)
    z, code, labels = synthetic.sample_celeba(np.random.RandomState(seed), Bg)
    full = [t.to(dev) for t in (imgs, z, code, labels)]
    mine = [parallel.shard(t) for t in full]

    step = CelebAStep(seed=seed, device=dev)
    parallel.attach(*step.optimizers())
    names = [["G." + n for n, _ in step.G.named_parameters()], ["D." + n for n, _ in step.D.named_parameters()]]
    names.append(names[0] + names[1])

    ref_rec, ref_losses, states = [], None, []
    if rank == 0:
        Fn.set_allreduce(None, 1)
        ref = CelebAStep(seed=seed, device=dev)
        parallel.detach(*ref.optimizers())

        def snap(i):
            states.append([{k: v.detach().clone() for k, v in net.state_dict().items()} for net in (ref.G, ref.D)])

        rl = ref(*full, record=ref_rec, after_phase=snap)
        ref_losses = torch.stack([rl["g_loss"], rl["d_loss"], rl["info_loss"]]).double()
        ref_bn = {k: v.clone() for k, v in ref.G.state_dict().items() if "running" in k}
        Fn.set_allreduce(dp.allreduce_sum_, world)
    bstates = []
    for i in range(2):
        per_net = []
        for ni, net in enumerate((step.G, step.D)):
            sd = {}
            for k, v in net.state_dict().items():
                t = states[i][ni][k].clone() if rank == 0 else torch.empty_like(v)
                dist.broadcast(t, 0)
                sd[k] = t
            per_net.append(sd)
        bstates.append(per_net)

    def load(i):
        step.G.load_state_dict(bstates[i][0])
        step.D.load_state_dict(bstates[i][1])

    red_log = []
    orig_reduce = dp.reduce

    def logging_reduce(opt):
        out = orig_reduce(opt)
        ps = [p for g in opt.param_groups for p in g["params"]]
        red_log.append([(out[p] / world).detach().clone() for p in ps])
        return out

    dp.reduce = logging_reduce
    try:
        losses = step(*mine, after_phase=load)
    finally:
        dp.reduce = orig_reduce
    loss_vec = torch.stack([losses["g_loss"], losses["d_loss"], losses["info_loss"]]).double()
    dist.all_reduce(loss_vec)
    loss_vec /= world
    torch.cuda.synchronize()
    if rank != 0:
        return None
    prec = os.environ.get("EADGAN_PRECISION", "bf16")
    tol_g, tol_l, tol_bn = (2e-4, 1e-5, 1e-4) if prec == "fp32" else (2e-2, 5e-3, 2e-2)
    rep = {"world": world, "global_batch": Bg, "precision": prec,
           "loss_max_abs_err": float((loss_vec - ref_losses).abs().max()), "phases": []}
    ok = rep["loss_max_abs_err"] <= tol_l * 10
    for ph in range(3):
        refs = dict(zip(names[ph], ref_rec[ph]["grads"]))
        worst, worst_name, num, den2 = 0.0, "", 0.0, 0.0
        for n, a, b in zip(names[ph], red_log[ph], ref_rec[ph]["grads"]):
            num += float((a - b).double().pow(2).sum())
            den2 += float(b.double().pow(2).sum())
            scale = float(refs[_ZERO_GRAD.get(n, n)].abs().max())
            if b.numel() < 16 and n.endswith(".bias"):        # [3]-element bias: one more column of its layer's gradient
                scale = max(scale, float(refs.get(n[:-5] + ".weight", b).abs().max()))
            e = float((a - b).abs().max()) / max(scale, 1e-30)
            if e > worst:
                worst, worst_name = e, n
        l2 = (num / max(den2, 1e-300)) ** 0.5
        rep["phases"].append({"worst_tensor_err": worst, "worst_tensor": worst_name, "l2_rel_all": l2})
        ok &= worst <= tol_g
    bn = 0.0
    sd = step.G.state_dict()
    for k, v in ref_bn.items():
        scale = float(v.abs().max()) if "var" in k else float(ref_bn[k.replace("mean", "var")].sqrt().max())
        bn = max(bn, float((sd[k] - v).abs().max()) / max(scale, 1e-12))
    rep["bn_running_stats_err"] = bn
    ok &= bn <= tol_bn
    rep["bounds"] = {"grad_tensor_max": tol_g, "loss_abs": tol_l * 10, "bn": tol_bn}
    rep["pass"] = bool(ok)
    return rep


