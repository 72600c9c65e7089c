"""GPU micro-benchmark: tcgen05 GEMM mainloop rate vs BLOCK_N / MT with L2-resident operands -> clocks per UMMA.
Run once per configuration: EADGAN_TC_BN=<bn> EADGAN_TC_MT=<mt> python tools/bench_gemm.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eadgan_b200 import tc
dev = torch.device("cuda:0")
reps = 20
M, K, N = 148 * 128, 1024, 2048
a = torch.randn(M, K, device=dev).bfloat16()
b = torch.randn(N, K, device=dev).bfloat16()
for _ in range(3):
    tc.gemm(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    tc.gemm(a, b)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
bn = int(os.environ.get("EADGAN_TC_BN", "256")); mt = int(os.environ.get("EADGAN_TC_MT", "1"))
n_umma = (M // 128) * (N // bn) * (K // 16) / 148   # per SM
print(f"BN={bn} MT={mt}: {ms:7.4f} ms {2.0*M*N*K/ms/1e9:8.1f} TF/s  -> {ms*1e-3*1.9e9/n_umma:6.1f} clk/UMMA(128x{bn}x16) at 1.9 GHz", flush=True)
