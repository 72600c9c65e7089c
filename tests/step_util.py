"""Shared helper of the step-level GPU tests and tools/diag_step.py: run the oracle and OUR CelebA step on
the same seeded inputs and weights, restarting every phase of ours from the ORACLE's post-phase state so
that Adam's sign(g) noise of one phase does not contaminate the next (SURVEY.md section 7.3-1 iv)."""
import os

import numpy as np
import torch


def tensor_err(a, b, floor=0.0):
    """tensor-normalised max error max|a-b| / max(max|b|, floor)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), floor, 1e-300)


def l2_err(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def cosine(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def run_pair(dev, B, precision, oracle_dtype=torch.float64, seed=0, sync=True):
    """-> (oracle record, our per-phase record, our losses, oracle state, our step object)."""
    from eadgan_b200.steps.celeba import CelebAStep
    from oracle import torch_oracle as O
    os.environ["EADGAN_PRECISION"] = precision
    imgs = O.synth_celeba_images(B, seed).to(dev)
    draws = O.sample_celeba(np.random.RandomState(seed), B)
    st = O.build_celeba(seed=seed, device=dev, dtype=oracle_dtype)
    ref = O.step_celeba(st, imgs.to(oracle_dtype), draws)
    ours = CelebAStep(seed=seed, device=dev)

    def after_phase(i):
        if not sync:
            return
        snap = ref["phases"][i]["state_after"]
        ours.G.load_state_dict({k: v.to(torch.float32) if v.is_floating_point() else v for k, v in snap["G"].items()})
        ours.D.load_state_dict({k: v.to(torch.float32) if v.is_floating_point() else v for k, v in snap["D"].items()})

    rec = []
    losses = ours(imgs, draws["z"].to(dev), draws["code"].to(dev), draws["labels"].to(dev), record=rec,
                  after_phase=after_phase)
    return ref, rec, {k: float(v) for k, v in losses.items()}, st, ours


def grad_names(step):
    g = [n for n, _ in step.G.named_parameters()]
    d = [n for n, _ in step.D.named_parameters()]
    return [["G." + n for n in g], ["D." + n for n in d], ["G." + n for n in g] + ["D." + n for n in d]]


# conv biases directly in front of a train-mode BatchNorm: their gradient is mathematically zero (the
# reference's own value is fp32 summation noise), so they are normalised by the sibling weight gradient
ZERO_GRAD = {"G.conv_blocks.1.bias": "G.conv_blocks.1.weight", "G.conv_blocks.4.bias": "G.conv_blocks.4.weight",
             "G.conv_blocks.7.bias": "G.conv_blocks.7.weight"}


def phase_errors(names, ours_grads, ref_grads, zero_grad=None):
    """-> {name: (max_err, l2_err, cosine)}.  Zero-gradient tensors use their sibling weight's scale
    (l2 / cosine = None); so do tensors with fewer than 16 elements (the [3] bias of G's last layer is
    a sum of 4096*B random-sign terms: ill-conditioned as a 3-vector, well-defined as one more column
    of the layer's [dW | db] gradient)."""
    refs = dict(zip(names, ref_grads))
    out = {}
    for n, a, b in zip(names, ours_grads, ref_grads):
        zg = ZERO_GRAD if zero_grad is None else zero_grad
        sib = zg.get(n)
        if sib is None and b.numel() < 16 and n.endswith(".bias"):
            sib = n[:-5] + ".weight" if n[:-5] + ".weight" in refs else n[:-5] + ".weight_orig"
        if sib is not None:
            floor = refs[sib].abs().max().item()
            out[n] = (tensor_err(a, b, floor), None, None if n in zg else "small")
        else:
            out[n] = (tensor_err(a, b), l2_err(a, b), cosine(a, b))
    return out


# ---- dSprites -------------------------------------------------------------------------------------------
def run_pair_dsprites(dev, B, precision, oracle_dtype=torch.float64, seed=0, sync=True):
    """the dSprites stage-2 step (dSprites/rp.py): oracle vs OUR step; phase info of ours restarts from the
    oracle's post-phase-D state."""
    from eadgan_b200.steps.dsprites import DSpritesStep
    from oracle import torch_oracle as O
    os.environ["EADGAN_PRECISION"] = precision
    imgs = O.synth_dsprites_images(B, seed).to(dev)
    draws = O.sample_dsprites(np.random.RandomState(seed), B)
    st = O.build_dsprites(seed=seed, device=dev, dtype=oracle_dtype)
    ref = O.step_dsprites(st, imgs, draws)
    ours = DSpritesStep(seed=seed, device=dev, pxy_state=O.dsprites_pxy_state(seed))

    def after_phase(i):
        if sync and i == 0:
            snap = ref["phases"][0]["state_after"]
            for key, net in (("G", ours.G), ("D", ours.D), ("E", ours.E)):
                net.load_state_dict({k: v.to(torch.float32) if v.is_floating_point() else v for k, v in snap[key].items()})

    rec = []
    losses = ours(imgs, draws["code_d"].to(dev), draws["labels_d"].to(dev), draws["code_info"].to(dev),
                  draws["labels_info"].to(dev), record=rec, after_phase=after_phase)
    return ref, rec, {k: float(v) for k, v in losses.items()}, st, ours


def dsprites_grad_names(step):
    d = ["D." + n for n, _ in step.D.named_parameters()]
    g = ["G." + n for n, _ in step.G.named_parameters()]
    e = ["E." + n for n, _ in step.E.named_parameters()]
    return [d, g + e]


DSPRITES_ZERO_GRAD = {"G.conv_block.0.bias": "G.conv_block.0.weight", "G.conv_block.3.bias": "G.conv_block.3.weight",
                      "G.conv_block.6.bias": "G.conv_block.6.weight"}


# ---- colored dSprites ------------------------------------------------------------------------------------
def run_pair_colored(dev, B, precision, oracle_dtype=torch.float64, seed=0, sync=True):
    """the colored-dSprites stage-2 step (colored_dSprites/rp_color.py): oracle vs OUR step."""
    from eadgan_b200.steps.colored import ColoredDSpritesStep
    from oracle import torch_oracle as O
    os.environ["EADGAN_PRECISION"] = precision
    imgs = O.synth_dsprites_images(B, seed).to(dev)
    draws = O.sample_colored(np.random.RandomState(seed), B)
    st = O.build_dsprites(seed=seed, device=dev, dtype=oracle_dtype, colored=True)
    ref = O.step_colored(st, imgs, draws)
    ours = ColoredDSpritesStep(seed=seed, device=dev, pxy_state=O.dsprites_pxy_state(seed, colored=True))

    def after_phase(i):
        if sync and i == 0:
            snap = ref["phases"][0]["state_after"]
            for key, net in (("G", ours.G), ("D", ours.D), ("E", ours.E)):
                net.load_state_dict({k: v.to(torch.float32) if v.is_floating_point() else v for k, v in snap[key].items()})

    rec = []
    losses = ours(imgs, draws["color"].to(dev), draws["code_d"].to(dev), draws["labels_d"].to(dev),
                  draws["code_info"].to(dev), draws["labels_info"].to(dev), record=rec, after_phase=after_phase)
    return ref, rec, {k: float(v) for k, v in losses.items()}, st, ours


# ---- MNIST -------------------------------------------------------------------------------------------------
def run_pair_mnist(dev, B, precision, oracle_dtype=torch.float64, seed=0, sync=True):
    """the MNIST step (MNIST/EAD-GAN_rpqmnxy.py): oracle vs OUR step, phases restarted from the oracle's state."""
    from eadgan_b200.steps.mnist import MnistStep
    from oracle import torch_oracle as O
    os.environ["EADGAN_PRECISION"] = precision
    imgs = O.synth_mnist_images(B, seed).to(dev)
    draws = O.sample_mnist(np.random.RandomState(seed), B)
    st = O.build_mnist(seed=seed, device=dev, dtype=oracle_dtype)
    ref = O.step_mnist(st, imgs, draws)
    ours = MnistStep(seed=seed, device=dev, approximator_state=O.mnist_approximator_state(seed))

    def after_phase(i):
        if not sync:
            return
        snap = ref["phases"][i]["state_after"]
        for key, net in (("G", ours.G), ("D", ours.D), ("E", ours.E)):
            net.load_state_dict({k: v.to(torch.float32) if v.is_floating_point() else v for k, v in snap[key].items()})

    rec = []
    losses = ours(imgs, draws["z"].to(dev), draws["code"].to(dev), draws["labels"].to(dev), record=rec,
                  after_phase=after_phase)
    return ref, rec, {k: float(v) for k, v in losses.items()}, st, ours


def mnist_grad_names(step):
    g = ["G." + n for n, _ in step.G.named_parameters()]
    d = ["D." + n for n, _ in step.D.named_parameters()]
    e = ["E." + n for n, _ in step.E.named_parameters()]
    return [g, d, g + e]


# conv biases in front of a train-mode BatchNorm (G: conv_blocks.2 -> BN .3, conv_blocks.6 -> BN .7; the l1 Linear
# feeds BatchNorm .0 directly): mathematically zero gradients
MNIST_ZERO_GRAD = {"G.conv_blocks.2.bias": "G.conv_blocks.2.weight", "G.conv_blocks.6.bias": "G.conv_blocks.6.weight",
                   "G.l1.0.bias": "G.l1.0.weight"}
