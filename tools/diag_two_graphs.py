import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eadgan_b200 import tc
from eadgan_b200.graph import GraphedStep
from eadgan_b200.steps.celeba import CelebAStep
from oracle.torch_oracle import sample_celeba, synth_celeba_images
cuda = torch.device("cuda:0")
B = 16
mode = sys.argv[1]
if mode == "nopool":
    tc._pool_enabled = False
d = sample_celeba(np.random.RandomState(20), B)
dev0 = [t.to(cuda) for t in (synth_celeba_images(B, 0), d["z"], d["code"], d["labels"])]
A = GraphedStep(CelebAStep(seed=0, device=cuda), dev0, warmup=1)
torch.cuda.synchronize(); print("A built", flush=True)
if mode == "emptycache":
    import gc; gc.collect(); torch.cuda.empty_cache()
    print("A:", {k: float(v) for k, v in A(*dev0).items()}, flush=True)
    sys.exit(0)
if mode == "noinvalidate":
    tc.invalidate_caches = lambda: None
if mode == "eager_only":
    s2 = CelebAStep(seed=0, device=cuda)
    s2(*dev0); torch.cuda.synchronize()
    print("A after eager step of another model:", {k: float(v) for k, v in A(*dev0).items()}, flush=True)
    sys.exit(0)
Bg = GraphedStep(CelebAStep(seed=0, device=cuda), dev0, warmup=1)
torch.cuda.synchronize(); print("B built", flush=True)
print("A:", {k: float(v) for k, v in A(*dev0).items()}, flush=True)
print("B:", {k: float(v) for k, v in Bg(*dev0).items()}, flush=True)
