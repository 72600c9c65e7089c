"""GPU micro-benchmark of the thin image-layer kernels at B = 1024 (CUDA events): D's Conv2d(3,128) forward
(channel-major tcgen05 kernel), its weight gradient and input gradient, the row-expansion pass.
usage: bench_thin.py [B] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eadgan_b200 import tc  # noqa: E402
from eadgan_b200._lib import ACT_LRELU  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


img = torch.rand(B, 3, 64, 64, device=dev) * 2 - 1
w = torch.randn(128, 3, 4, 4, device=dev) * 0.1
b = torch.randn(128, device=dev)
sigma = torch.tensor([1.3], device=dev)
r = tc.thin_expand(img)
wf, wd = tc.thin_pack_w(w, "fprop"), tc.thin_pack_w(w, "dgrad")
out = tc.alloc_padded(B, 32, 32, 128, dev)
dy = tc.alloc_padded(B, 32, 32, 128, dev)
tc.interior(dy).normal_()
out_bytes = B * 32 * 32 * 128 * 2
cases = {
    "thin_expand": (lambda: tc.thin_expand(img), B * 3 * 64 * 64 * 4 + r.numel() * 2),
    "thin_fprop+bias+lrelu+sigma": (lambda: tc.thin_fprop(r, wf, b, 3, 128, ACT_LRELU, 0.1, sigma=sigma, out=out), r.numel() * 2 + out_bytes),
    "thin_wgrad": (lambda: tc.thin_wgrad(r, dy, 3), r.numel() * 2 + out_bytes),
    "thin_dgrad": (lambda: tc.thin_dgrad(dy, wd, None, 3, sigma=sigma), out_bytes + B * 3 * 64 * 64 * 4),
}
for name, (fn, nbytes) in cases.items():
    ms = timeit(fn)
    print(f"B={B} {name:30s} {ms:8.4f} ms  {nbytes / ms / 1e6:8.0f} GB/s (algorithmic bytes {nbytes / 1e6:.0f} MB)", flush=True)
