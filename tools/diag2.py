"""GPU diagnostic: conv-bias gradient in front of a train-mode BatchNorm (fp32 per-op path)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import eadgan_b200.nn as enn
from eadgan_b200 import functional as Fn
dev = torch.device("cuda:0")
os.environ["EADGAN_PRECISION"] = "fp32"
torch.manual_seed(0)
B = 8
ours = enn.Sequential(enn.ConvTranspose2d(256, 128, 4, 2, 1), enn.BatchNorm2d(128), enn.ReLU()).to(dev)
ref = torch.nn.Sequential(torch.nn.ConvTranspose2d(256, 128, 4, 2, 1), torch.nn.BatchNorm2d(128), torch.nn.ReLU()).to(dev)
ref.load_state_dict(ours.state_dict())
x = torch.randn(B, 256, 16, 16, device=dev)
# manual per-module run keeping intermediates
h0 = ours[0](x); h0.retain_grad()
y = ours[1](h0, act=ours[2].act())
r0 = ref[0](x); r0.retain_grad()
ry = ref[2](ref[1](r0))
go = torch.randn_like(ry)
y.backward(go); ry.backward(go)
print("y err", (y - ry).abs().max().item())
print("dh0 err", (h0.grad - r0.grad).abs().max().item(), "max", r0.grad.abs().max().item())
print("sum dh0 ours (torch sum)", h0.grad.sum((0, 2, 3))[:6].tolist())
print("sum dh0 ref  (torch sum)", r0.grad.sum((0, 2, 3))[:6].tolist())
print("channel_sum(ours dh0)   ", Fn.channel_sum(h0.grad)[:6].tolist())
print("bias grad ours", ours[0].bias.grad[:6].tolist())
print("bias grad ref ", ref[0].bias.grad[:6].tolist())
print("wgrad err", (ours[0].weight.grad - ref[0].weight.grad).abs().max().item() / ref[0].weight.grad.abs().max().item())
