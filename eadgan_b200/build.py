"""Build libeadgan.so in-tree: ``python -m eadgan_b200.build`` (nvcc, sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = ["simt_conv.cu", "bn.cu", "elementwise.cu", "spectral_norm.cu", "loss_adam.cu", "tc_conv.cu", "glue.cu", "sample.cu"]


def build(force=False, verbose=False):
    csrc = os.path.join(HERE, "csrc")
    out = os.path.join(HERE, "lib", "libeadgan.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(csrc, s) for s in SRC]
    deps = srcs + [os.path.join(csrc, "common.cuh"), os.path.join(HERE, "..", "include", "eadgan.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(HERE, "..", "include"),
           "-diag-suppress", "177", *srcs, "-o", out]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.run(cmd, check=True)
    return out


def build_experimental(force=False):
    """csrc/experimental/*.cu are NOT part of libeadgan.so (nothing in the product loads them); they are compiled
    here only so that the build check covers every CUDA source in the tree."""
    csrc = os.path.join(HERE, "csrc", "experimental")
    outs = []
    for name in sorted(os.listdir(csrc)) if os.path.isdir(csrc) else []:
        if not name.endswith(".cu"):
            continue
        src = os.path.join(csrc, name)
        out = os.path.join(HERE, "lib", "libeadgan_x_" + name[:-3] + ".so")
        if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                            "-Xcompiler", "-fPIC", "-shared", src, "-o", out], check=True)
        outs.append(out)
    return outs


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
