"""GPU micro-benchmark of the tcgen05 conv entry points on the CelebA layer geometries (CUDA events,
L2 flushed by the working set itself at B >= 512).  usage: bench_layers.py [B] [filter] [reps] [c]
(filter: substring of the case name, "=name" for an exact match; c: only the layer with that channel count)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eadgan_b200 import tc  # noqa: E402
from eadgan_b200._lib import ACT_LRELU, ACT_NONE  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
flt = sys.argv[2] if len(sys.argv) > 2 else ""
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
only_c = int(sys.argv[4]) if len(sys.argv) > 4 else 0
LAYERS = [(32, 64, 128), (128, 32, 256), (256, 16, 512), (512, 8, 1024)]  # (c, h, k) conv view


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


for c, h, k in LAYERS:
    if only_c and c != only_c:
        continue
    flops = 2.0 * B * (h // 2) ** 2 * k * c * 16
    xp = tc.alloc_padded(B, h, h, c, dev)
    tc.interior(xp).normal_()
    yp = tc.alloc_padded(B, h // 2, h // 2, k, dev)
    tc.interior(yp).normal_()
    w = torch.randn(k, c, 4, 4, device=dev) * 0.02
    bias_k, bias_c = torch.randn(k, device=dev), torch.randn(c, device=dev)
    wf, wd = tc.pack_w(w, None, "fprop"), tc.pack_w(w, None, "dgrad")
    out_s, out_b = torch.empty_like(yp), torch.empty_like(xp)
    stats = torch.zeros(2 * max(c, k), device=dev, dtype=torch.float64)
    cases = {
        "fprop": lambda: tc.fprop(xp, wf, None, k, out=out_s),
        "fprop+bias": lambda: tc.fprop(xp, wf, bias_k, k, out=out_s),
        "fprop+lrelu": lambda: tc.fprop(xp, wf, None, k, ACT_LRELU, 0.1, out=out_s),
        "fprop+bias+lrelu": lambda: tc.fprop(xp, wf, bias_k, k, ACT_LRELU, 0.1, out=out_s),
        "fprop+mask": lambda: tc.fprop(xp, wf, None, k, mask=yp, mask_mode=ACT_LRELU, slope=0.1, out=out_s),
        "dgrad": lambda: tc.dgrad(yp, wd, None, c, out=out_b),
        "dgrad+bias": lambda: tc.dgrad(yp, wd, bias_c, c, out=out_b),
        "dgrad+stats": lambda: tc.dgrad(yp, wd, None, c, stats=stats, out=out_b),
        "dgrad+bias+stats": lambda: tc.dgrad(yp, wd, bias_c, c, stats=stats, out=out_b),
        "dgrad+mask": lambda: tc.dgrad(yp, wd, None, c, mask=xp, mask_mode=ACT_LRELU, slope=0.1, out=out_b),
        "wgrad": lambda: tc.wgrad(xp, yp),
    }
    for name, fn in cases.items():
        if flt and (name != flt[1:] if flt.startswith("=") else flt not in name):
            continue
        ms = timeit(fn)
        print(f"B={B} c={c:4d} h={h:3d} k={k:4d} {name:18s} {ms:8.4f} ms  {flops / ms / 1e9:8.1f} TF/s", flush=True)

if only_c:
    sys.exit(0)
# calibration of this box: cuBLAS bf16 8192^3 (the MEASURED_PEAKS.json denominator), same process
a = torch.randn(8192, 8192, device=dev).bfloat16(); b = torch.randn(8192, 8192, device=dev).bfloat16()
print(f"calibration: torch.matmul bf16 8192^3 {2.0 * 8192 ** 3 / timeit(lambda: torch.matmul(a, b)) / 1e9:8.1f} TF/s")
