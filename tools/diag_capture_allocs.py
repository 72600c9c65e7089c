"""which allocations made DURING the whole-step capture come from the regular (non-graph) pool?"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eadgan_b200 import graph as G
from eadgan_b200.steps.celeba import CelebAStep
from oracle.torch_oracle import sample_celeba, synth_celeba_images
cuda = torch.device("cuda:0")
B = 16
d = sample_celeba(np.random.RandomState(20), B)
dev0 = [t.to(cuda) for t in (synth_celeba_images(B, 0), d["z"], d["code"], d["labels"])]
orig = torch.cuda.graph


class Spy(orig):
    def __enter__(self):
        r = super().__enter__()
        torch.cuda.memory._record_memory_history(max_entries=200000, stacks="python")
        return r

    def __exit__(self, *a):
        self.snap = torch.cuda.memory._snapshot()
        torch.cuda.memory._record_memory_history(enabled=None)
        Spy.last = self.snap
        return super().__exit__(*a)


torch.cuda.graph = Spy
A = G.GraphedStep(CelebAStep(seed=0, device=cuda), dev0, warmup=1)
snap = Spy.last
segs = [(s["address"], s["address"] + s["total_size"], tuple(s.get("segment_pool_id", (0, 0)))) for s in snap["segments"]]
def pool_of(addr):
    for a, b, p in segs:
        if a <= addr < b:
            return p
    return None
bad = {}
n = 0
for e in snap["device_traces"][0]:
    if e["action"] != "alloc":
        continue
    n += 1
    p = pool_of(e["addr"])
    if p == (0, 0) or p is None:
        fr = [f"{os.path.basename(f['filename'])}:{f['line']}:{f['name']}" for f in e.get("frames", []) if "eadgan" in f["filename"] or "steps" in f["filename"]][:4]
        key = " <- ".join(fr) or "(no eadgan frame)"
        bad.setdefault(key, [0, 0, p])
        bad[key][0] += 1; bad[key][1] += e["size"]
print("allocs during capture:", n, " from the regular pool:", sum(v[0] for v in bad.values()))
for k, v in sorted(bad.items(), key=lambda kv: -kv[1][1])[:25]:
    print(v, k)
