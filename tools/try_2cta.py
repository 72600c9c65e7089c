"""Round-2 starting point: build and try the EXPERIMENTAL cta_group::2 GEMM (eadgan_b200/csrc/experimental/
tc_gemm_2cta.cu; it assembles -- UTCHMMA.2CTA / UTMALDG.2D.2CTA / UTCBAR.2CTA.MULTICAST in SASS -- but has never run).

    gpurun --timeout 300 -- 'timeout 60 python tools/try_2cta.py'        # ALWAYS under a short timeout: a protocol
                                                                          # error in a 2-CTA kernel is a hang

Checks C = A . B^T against torch on a few shapes, then times 8192^3 next to the cta_group::1 kernel (eadgan_tc_gemm)
and cuBLAS.  Nothing in the product imports this."""
import ctypes as C
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SRC = os.path.join(ROOT, "eadgan_b200", "csrc", "experimental", "tc_gemm_2cta.cu")
OUT = os.path.join(ROOT, "eadgan_b200", "lib", "libeadgan_x_tc_gemm_2cta.so")   # same file eadgan_b200.build.build_experimental writes


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
                        "-std=c++17", "-Xcompiler", "-fPIC", "-shared", SRC, "-o", OUT], check=True)
    lib = C.CDLL(OUT)
    lib.eadgan_x_gemm_2cta.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


def main():
    lib = build()
    if not torch.cuda.is_available():
        print("built", OUT, "(no GPU here)")
        return
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def gemm(a, b):
        c = torch.empty((a.shape[0], b.shape[0]), device=dev, dtype=torch.float32)
        rc = lib.eadgan_x_gemm_2cta(a.data_ptr(), b.data_ptr(), c.data_ptr(), a.shape[0], b.shape[0], a.shape[1], st)
        if rc:
            raise RuntimeError(f"eadgan_x_gemm_2cta rc={rc}")
        return c

    torch.manual_seed(0)
    for m, n, k in ((256, 256, 64), (256, 256, 512), (512, 768, 1024), (300, 256, 256)):
        a = torch.randn(m, k, device=dev).bfloat16()
        b = torch.randn(n, k, device=dev).bfloat16()
        out = gemm(a, b)
        torch.cuda.synchronize()
        ref = a.float() @ b.float().t()
        err = float((out - ref).abs().max() / ref.abs().max())
        print(f"m={m} n={n} k={k}: rel err {err:.2e}", flush=True)
        assert err < 1e-4

    def timeit(fn, reps=10):
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

    n = 8192
    a = torch.randn(n, n, device=dev).bfloat16()
    b = torch.randn(n, n, device=dev).bfloat16()
    from eadgan_b200 import tc
    for name, fn in (("2-CTA (experimental, one tile per pair, non-persistent)", lambda: gemm(a, b)),
                     ("cta_group::1 persistent (eadgan_tc_gemm)", lambda: tc.gemm(a, b)),
                     ("cuBLAS", lambda: torch.matmul(a, b.t()))):
        ms = timeit(fn)
        print(f"{name:58s} {ms:8.3f} ms  {2.0 * n ** 3 / ms / 1e9:8.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
