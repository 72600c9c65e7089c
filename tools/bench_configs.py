"""Throughput of the other BASELINE configurations on one GPU (CUDA events, eager launch and whole-step CUDA
graph replay):  configs[1] dSprites rp.py batch 256, configs[2] colored rp_color.py batch 512 (its 1-GPU
equivalent), configs[0] MNIST batch 64 (BASELINE's CPU configuration, here for reference), stage-1 pxy batch 128.
usage: python tools/bench_configs.py [steps]   -> one JSON line per configuration"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eadgan_b200.graph import GraphedStep  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402  (synthetic-input generators only)

dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(name, make_step, inputs, batch):
    step = make_step()
    for _ in range(3):
        step(*inputs)
    eager_ms = timed(lambda: step(*inputs), steps)
    g = GraphedStep(make_step(), inputs, warmup=3)
    for _ in range(3):
        g(*inputs)
    graph_ms = timed(lambda: g(*inputs), steps)
    print(json.dumps({"config": name, "batch": batch, "eager_ms_per_step": eager_ms, "graph_ms_per_step": graph_ms,
                      "images_per_s_graph": batch / graph_ms * 1e3, "images_per_s_eager": batch / eager_ms * 1e3,
                      "kernels_per_step": g.kernels_per_replay, "precision": os.environ.get("EADGAN_PRECISION", "bf16")}),
          flush=True)


def main():
    from eadgan_b200.steps.colored import ColoredDSpritesStep
    from eadgan_b200.steps.dsprites import DSpritesStep
    from eadgan_b200.steps.mnist import MnistStep
    from eadgan_b200.steps.pxy import PxyStep
    rs = np.random.RandomState(0)
    B = 256
    d = O.sample_dsprites(rs, B)
    run("dsprites rp.py (BASELINE configs[1])",
        lambda: DSpritesStep(seed=0, device=dev, pxy_state=O.dsprites_pxy_state(0)),
        [O.synth_dsprites_images(B, 0).to(dev)] + [d[k].to(dev) for k in ("code_d", "labels_d", "code_info", "labels_info")], B)
    B = 512
    d = O.sample_colored(rs, B)
    run("colored rp_color.py (BASELINE configs[2], one GPU)",
        lambda: ColoredDSpritesStep(seed=0, device=dev, pxy_state=O.dsprites_pxy_state(0, colored=True)),
        [O.synth_dsprites_images(B, 0).to(dev)] + [d[k].to(dev) for k in ("color", "code_d", "labels_d", "code_info", "labels_info")], B)
    B = 64
    d = O.sample_mnist(rs, B)
    run("MNIST EAD-GAN_rpqmnxy.py (BASELINE configs[0] on the GPU)",
        lambda: MnistStep(seed=0, device=dev, approximator_state=O.mnist_approximator_state(0)),
        [O.synth_mnist_images(B, 0).to(dev), d["z"].to(dev), d["code"].to(dev), d["labels"].to(dev)], B)
    B = 128
    d = O.sample_pxy(rs, B)
    run("dSprites pxy.py stage 1", lambda: PxyStep(seed=0, device=dev),
        [O.synth_dsprites_images(B, 0).to(dev), d["code"].to(dev)], B)


if __name__ == "__main__":
    main()
