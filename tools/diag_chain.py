"""GPU diagnostic (not a test): mini-chains in bf16 vs stock torch fp32, per-tensor errors."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import eadgan_b200.nn as enn  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    d = b.abs().max().item()
    return (a - b).abs().max().item() / (d if d > 0 else 1.0)


def build(ns, spec, sn=False):
    layers = []
    for item in spec:
        kind = item[0]
        if kind == "conv":
            m = ns.Conv2d(*item[1:])
            layers.append((enn.spectral_norm if ns is enn else torch.nn.utils.spectral_norm)(m) if sn else m)
        elif kind == "convT":
            layers.append(ns.ConvTranspose2d(*item[1:]))
        elif kind == "bn":
            layers.append(ns.BatchNorm2d(item[1]))
        elif kind == "lrelu":
            layers.append(ns.LeakyReLU(item[1], inplace=True))
        elif kind == "relu":
            layers.append(ns.ReLU())
        elif kind == "tanh":
            layers.append(ns.Tanh())
    return ns.Sequential(*layers)


def run(name, spec, in_shape, sn=False, scale=1.0):
    torch.manual_seed(0)
    ours = build(enn, spec, sn).to(dev)
    ref = build(torch.nn, spec, sn).to(dev)
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(*in_shape, device=dev) * scale
    for prec in ("fp32", "bf16"):
        os.environ["EADGAN_PRECISION"] = prec
        ours.load_state_dict(ref.state_dict())
        ref2 = build(torch.nn, spec, sn).to(dev)
        ref2.load_state_dict(ref.state_dict())
        xo, xr = x.clone().requires_grad_(), x.clone().requires_grad_()
        yo, yr = ours(xo), ref2(xr)
        torch.manual_seed(1)
        go = torch.randn_like(yr)
        po, pr = [xo] + list(ours.parameters()), [xr] + list(ref2.parameters())
        gso, gsr = torch.autograd.grad(yo, po, go), torch.autograd.grad(yr, pr, go)
        names = ["x"] + [n for n, _ in ours.named_parameters()]
        errs = {n: (rel(a, b) if b.abs().max() > 1e-6 else float((a - b).abs().max())) for n, a, b in zip(names, gso, gsr)}
        print(f"[{name}] {prec}: out {rel(yo, yr):.2e} | " + " ".join(f"{n}:{e:.1e}" for n, e in errs.items()), flush=True)
        if prec == "bf16":
            import bf16_emul
            ref3 = build(torch.nn, spec, sn).to(dev)
            ref3.load_state_dict(ref.state_dict())
            xe = x.clone().requires_grad_()
            ye = bf16_emul.emulate(ours, ref3, xe)
            gse = torch.autograd.grad(ye, [xe] + list(ref3.parameters()), go)
            errs = {n: (rel(a, b) if b.abs().max() > 1e-6 else float((a - b).abs().max())) for n, a, b in zip(names, gso, gse)}
            print(f"[{name}] bf16 vs EMUL: out {rel(yo, ye):.2e} | " + " ".join(f"{n}:{e:.1e}" for n, e in errs.items()), flush=True)


B = 8
run("conv128-256 single", [("conv", 128, 256, 4, 2, 1)], (B, 128, 32, 32))
run("conv128-256+lrelu", [("conv", 128, 256, 4, 2, 1), ("lrelu", 0.1)], (B, 128, 32, 32))
run("2conv+lrelu", [("conv", 128, 256, 4, 2, 1), ("lrelu", 0.1), ("conv", 256, 512, 4, 2, 1), ("lrelu", 0.1)], (B, 128, 32, 32))
run("3conv+lrelu+head", [("conv", 128, 256, 4, 2, 1), ("lrelu", 0.1), ("conv", 256, 512, 4, 2, 1), ("lrelu", 0.1),
                         ("conv", 512, 1024, 4, 2, 1), ("lrelu", 0.1), ("conv", 1024, 19, 4, 1, 0)], (B, 128, 32, 32))
run("D full", [("conv", 3, 128, 4, 2, 1), ("lrelu", 0.1), ("conv", 128, 256, 4, 2, 1), ("lrelu", 0.1),
               ("conv", 256, 512, 4, 2, 1), ("lrelu", 0.1), ("conv", 512, 1024, 4, 2, 1), ("lrelu", 0.1),
               ("conv", 1024, 19, 4, 1, 0)], (B, 3, 64, 64))
run("D full SN", [("conv", 3, 128, 4, 2, 1), ("lrelu", 0.1), ("conv", 128, 256, 4, 2, 1), ("lrelu", 0.1),
                  ("conv", 256, 512, 4, 2, 1), ("lrelu", 0.1), ("conv", 512, 1024, 4, 2, 1), ("lrelu", 0.1),
                  ("conv", 1024, 19, 4, 1, 0)], (B, 3, 64, 64), sn=True)
run("convT single", [("convT", 256, 128, 4, 2, 1)], (B, 256, 16, 16))
run("convT+bn+relu", [("convT", 256, 128, 4, 2, 1), ("bn", 128), ("relu",)], (B, 256, 16, 16))
run("convT+bn+relu x2 + tanh", [("convT", 512, 256, 4, 2, 1), ("bn", 256), ("relu",), ("convT", 256, 128, 4, 2, 1),
                                ("bn", 128), ("relu",), ("convT", 128, 3, 4, 2, 1), ("tanh",)], (B, 512, 8, 8))
run("G full", [("convT", 218, 1024, 4, 1, 0), ("convT", 1024, 512, 4, 2, 1), ("bn", 512), ("relu",),
               ("convT", 512, 256, 4, 2, 1), ("bn", 256), ("relu",), ("convT", 256, 128, 4, 2, 1), ("bn", 128),
               ("relu",), ("convT", 128, 3, 4, 2, 1), ("tanh",)], (B, 218, 1, 1))
