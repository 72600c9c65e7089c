"""per-entry-point time table of one eager step of a configuration: profile_config.py dsprites|colored|mnist [B]"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eadgan_b200 import _lib
from oracle import torch_oracle as O
dev = torch.device("cuda:0")
cfg = sys.argv[1]
rs = np.random.RandomState(0)
if cfg == "dsprites":
    from eadgan_b200.steps.dsprites import DSpritesStep
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    d = O.sample_dsprites(rs, B)
    step = DSpritesStep(seed=0, device=dev, pxy_state=O.dsprites_pxy_state(0))
    inputs = [O.synth_dsprites_images(B, 0).to(dev)] + [d[k].to(dev) for k in ("code_d", "labels_d", "code_info", "labels_info")]
elif cfg == "colored":
    from eadgan_b200.steps.colored import ColoredDSpritesStep
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    d = O.sample_colored(rs, B)
    step = ColoredDSpritesStep(seed=0, device=dev, pxy_state=O.dsprites_pxy_state(0, colored=True))
    inputs = [O.synth_dsprites_images(B, 0).to(dev)] + [d[k].to(dev) for k in ("color", "code_d", "labels_d", "code_info", "labels_info")]
else:
    from eadgan_b200.steps.mnist import MnistStep
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    d = O.sample_mnist(rs, B)
    step = MnistStep(seed=0, device=dev, approximator_state=O.mnist_approximator_state(0))
    inputs = [O.synth_mnist_images(B, 0).to(dev), d["z"].to(dev), d["code"].to(dev), d["labels"].to(dev)]
for _ in range(3):
    step(*inputs)
_lib.profile_start()
step(*inputs)
prof = _lib.profile_stop()
tot = sum(v["ms"] for v in prof.values())
print(f"{cfg} B={B}: sum of entry-point times {tot:.3f} ms")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:25]:
    print(f"  {k:56s} calls={v['calls']:4d} ms={v['ms']:.3f}")
