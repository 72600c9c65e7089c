"""Flat training script for tests/test_dp_gpu.py::test_run_py_under_torchrun: stock torch.nn names, one Adam, every
rank draws DIFFERENT data (seeded by RANK), identical initial weights (seeded the same).  Under
``torchrun -m eadgan_b200.run`` the optimiser all-reduces its gradients and BatchNorm statistics are synchronised
without the script knowing about data parallelism, so the replicas must stay bit-identical."""
import json
import os

import torch
import torch.nn as nn

rank = int(os.environ.get("RANK", "0"))
torch.manual_seed(0)
net = nn.Sequential(nn.ConvTranspose2d(16, 64, 4, 1, 0), nn.ConvTranspose2d(64, 32, 4, 2, 1), nn.BatchNorm2d(32), nn.ReLU(),
                    nn.ConvTranspose2d(32, 3, 4, 2, 1), nn.Tanh()).cuda()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.5, 0.999))
loss_fn = torch.nn.MSELoss().cuda()
g = torch.Generator().manual_seed(100 + rank)
for it in range(4):
    z = torch.randn(8, 16, 1, 1, generator=g).cuda()
    target = (torch.rand(8, 3, 16, 16, generator=g) * 2 - 1).cuda()
    opt.zero_grad()
    loss = loss_fn(net(z), target)
    loss.backward()
    opt.step()
sd = net.state_dict()
print(json.dumps({"rank": rank, "loss": loss.item(), "opt": type(opt).__module__,
                  "checksum": [float(v.double().sum()) for v in sd.values() if v.is_floating_point()]}), flush=True)
