"""GPU diagnostic: per-phase gradients of the CelebA step -- ours (fp32 / bf16) and the torch fp32 oracle,
each against the torch fp64 oracle, phases restarted from the oracle's state (tests/step_util.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import step_util as U  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402
import numpy as np  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
precs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fp32", "bf16"]

imgs = O.synth_celeba_images(B, seed).to(dev)
draws = O.sample_celeba(np.random.RandomState(seed), B)
r32 = O.step_celeba(O.build_celeba(seed, device=dev), imgs, draws)
for prec in precs:
    ref, rec, losses, st, ours = U.run_pair(dev, B, prec, seed=seed)
    names = U.grad_names(ours)
    print(f"B={B} seed={seed} losses fp64 {ref['losses']}\n   ref32 {r32['losses']}\n   ours-{prec} {losses}")
    for ph in range(3):
        eo = U.phase_errors(names[ph], rec[ph]["grads"], ref["phases"][ph]["grads"])
        er = U.phase_errors(names[ph], r32["phases"][ph]["grads"], ref["phases"][ph]["grads"]) if ph == 0 else None
        mx = [v[0] for v in eo.values()]
        l2 = [v[1] for v in eo.values() if v[1] is not None]
        cs = [v[2] for v in eo.values() if isinstance(v[2], float)]
        print(f"  phase {ph} ours-{prec} vs fp64: max-err max {max(mx):.2e} median {sorted(mx)[len(mx)//2]:.2e} | "
              f"L2 max {max(l2):.2e} median {sorted(l2)[len(l2)//2]:.2e} | cos min {min(cs):.5f}")
        if er is not None:
            mr = [v[0] for v in er.values()]
            print(f"          ref32 vs fp64: max-err max {max(mr):.2e} median {sorted(mr)[len(mr)//2]:.2e}")
        if "-v" in sys.argv:
            for (n, v), gr in zip(eo.items(), ref["phases"][ph]["grads"]):
                print(f"      {n:32s} |g|max {gr.abs().max().item():.2e} max {v[0]:.2e}" + ("" if v[1] is None else f" l2 {v[1]:.2e} cos {v[2]:.5f}"))
