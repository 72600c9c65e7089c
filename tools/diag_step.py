"""GPU diagnostic: phase-G gradients of the CelebA step -- ours (fp32 / bf16) and the torch fp32
oracle, each against the torch fp64 oracle (SURVEY.md section 7.3-1 referee protocol)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from eadgan_b200.steps.celeba import CelebAStep  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    d = b.abs().max().item()
    return (a - b).abs().max().item() / d if d > 1e-6 else (a - b).abs().max().item()


imgs = O.synth_celeba_images(B, 0).to(dev)
draws = O.sample_celeba(np.random.RandomState(0), B)
r64 = O.step_celeba(O.build_celeba(0, device=dev, dtype=torch.float64), imgs.double(), draws)
r32 = O.step_celeba(O.build_celeba(0, device=dev), imgs, draws)
print("losses fp64", r64["losses"])
print("losses ref32", r32["losses"])
for prec in ("fp32", "bf16"):
    os.environ["EADGAN_PRECISION"] = prec
    rec = []
    out = CelebAStep(seed=0, device=dev)(imgs, draws["z"].to(dev), draws["code"].to(dev), draws["labels"].to(dev), record=rec)
    print(f"losses ours {prec}", {k: float(v) for k, v in out.items()})
    for ph in range(3):
        eo = [rel(a, b) for a, b in zip(rec[ph]["grads"], r64["phases"][ph]["grads"])]
        er = [rel(a, b) for a, b in zip(r32["phases"][ph]["grads"], r64["phases"][ph]["grads"])]
        print(f"  phase {ph} ours-{prec} vs fp64: max {max(eo):.2e} median {sorted(eo)[len(eo)//2]:.2e} | ref32 vs fp64: max {max(er):.2e} median {sorted(er)[len(er)//2]:.2e}")
        if ph == 0:
            print("    ours:", " ".join(f"{e:.1e}" for e in eo))
            print("    ref :", " ".join(f"{e:.1e}" for e in er))
