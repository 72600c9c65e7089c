"""-m gpu: the honest competitor on the same box: the ORACLE -- the reference's own stock
torch.nn modules and step -- run on the B200 through cuDNN / cuBLAS, timed next to our CUDA-graph step on identical
inputs.  Writes gpurun_out/stock_torch_vs_ours.json and asserts that the product path is not slower than the stock
one in its fastest mode (TF32)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def test_stock_torch_step_vs_ours(cuda):
    from eadgan_b200.graph import GraphedStep
    from eadgan_b200.steps.celeba import CelebAStep
    from oracle import torch_oracle as O
    B = 512
    imgs = O.synth_celeba_images(B, 0).to(cuda)
    d = O.sample_celeba(np.random.RandomState(0), B)
    out = {"batch": B, "config": "CelebA EAD-GAN_celebA 64x64 step (3 phases, 3 Adams)"}
    st = O.build_celeba(seed=0, device=cuda)
    # bf16 autocast is not an option for the unmodified step: torch refuses BCELoss on a sigmoid output under
    # autocast ("unsafe to autocast"), so TF32 is the fastest stock mode that runs the reference as written
    out["stock_bf16_autocast"] = "rejected by torch: nn.BCELoss is unsafe to autocast"
    for name, tf32 in (("stock_fp32", False), ("stock_tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32

        def step():
            O.step_celeba(st, imgs, d, record=False)
        for _ in range(2):
            step()
        out[name + "_ms"] = _time(step, 3)
        out[name + "_images_per_s"] = B / out[name + "_ms"] * 1e3
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    os.environ["EADGAN_PRECISION"] = "bf16"
    dev_in = [imgs, d["z"].to(cuda), d["code"].to(cuda), d["labels"].to(cuda)]
    g = GraphedStep(CelebAStep(seed=0, device=cuda), dev_in, warmup=2)
    for _ in range(2):
        g(*dev_in)
    out["ours_bf16_graph_ms"] = _time(lambda: g(*dev_in), 5)
    out["ours_bf16_graph_images_per_s"] = B / out["ours_bf16_graph_ms"] * 1e3
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "stock_torch_vs_ours.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
    assert out["ours_bf16_graph_ms"] <= out["stock_tf32_ms"], out
