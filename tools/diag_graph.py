import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["EADGAN_PRECISION"] = sys.argv[1] if len(sys.argv) > 1 else "bf16"
from eadgan_b200.graph import GraphedStep
from eadgan_b200.steps.celeba import CelebAStep
from oracle.torch_oracle import sample_celeba, synth_celeba_images
dev = torch.device("cuda:0")
B = 16
def batch(i):
    d = sample_celeba(np.random.RandomState(10 + i), B)
    return [synth_celeba_images(B, i).to(dev), d["z"].to(dev), d["code"].to(dev), d["labels"].to(dev)]
def fmt(o): return " ".join(f"{float(v):.6f}" for v in o.values())
seq = [0, 0, 1, 2, 3]
a, b = CelebAStep(seed=0, device=dev), CelebAStep(seed=0, device=dev)
print("eager A vs eager B (default stream)")
for i in seq:
    print("  A", fmt(a(*batch(i))), "| B", fmt(b(*batch(i))))
c = CelebAStep(seed=0, device=dev)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
print("eager C on a side stream")
with torch.cuda.stream(side):
    for i in seq:
        print("  C", fmt(c(*batch(i))))
torch.cuda.synchronize()
g = CelebAStep(seed=0, device=dev)
gs = GraphedStep(g, batch(0), warmup=2)
print("graphed (2 eager warm-up steps on batch 0 inside), then batches 1,2,3")
for i in (1, 2, 3):
    print("  G", fmt(gs(*batch(i))))
