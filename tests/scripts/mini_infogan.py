"""A flat training script written the way the reference's scripts are (module-level argparse, models built from
stock ``torch.nn`` names with ``opt`` read inside the constructors, ``spectral_norm`` imported from torch, losses and
two Adams, a training loop at import time) -- but small, self-contained and on synthetic data, so that it can run on
the GPU box where /root/reference does not exist.  tests/test_run_gpu.py runs it once with stock PyTorch and once,
UNCHANGED, under ``python -m eadgan_b200.run`` and compares the printed losses."""
import argparse
import json

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm

parser = argparse.ArgumentParser()
parser.add_argument("--n_iter", type=int, default=3)
parser.add_argument("--batch_size", type=int, default=16)
parser.add_argument("--latent_dim", type=int, default=20)
parser.add_argument("--code_dim", type=int, default=4)
parser.add_argument("--n_classes", type=int, default=5)
parser.add_argument("--seed", type=int, default=0)
opt = parser.parse_args()

cuda = torch.cuda.is_available()
torch.manual_seed(opt.seed)
np.random.seed(opt.seed)


class Generator(nn.Module):
    def __init__(self):
        super(Generator, self).__init__()
        in_dim = opt.latent_dim + opt.n_classes + opt.code_dim
        self.conv_blocks = nn.Sequential(
            nn.ConvTranspose2d(in_dim, 128, 4, 1, 0),
            nn.ConvTranspose2d(128, 64, 4, stride=2, padding=1), nn.BatchNorm2d(64), nn.ReLU(),
            nn.ConvTranspose2d(64, 64, 4, stride=2, padding=1), nn.BatchNorm2d(64), nn.ReLU(),
            nn.ConvTranspose2d(64, 32, 4, stride=2, padding=1), nn.BatchNorm2d(32), nn.ReLU(),
            nn.ConvTranspose2d(32, 3, 4, stride=2, padding=1), nn.Tanh(),
        )

    def forward(self, noise, labels, code):
        g_in = torch.cat((noise, labels, code), -1)
        return self.conv_blocks(g_in.view(g_in.size(0), g_in.size(1), 1, 1))


class Discriminator(nn.Module):
    def __init__(self):
        super(Discriminator, self).__init__()
        self.main = nn.Sequential(
            spectral_norm(nn.Conv2d(3, 32, 4, 2, 1)), nn.LeakyReLU(0.1, inplace=True),
            spectral_norm(nn.Conv2d(32, 64, 4, 2, 1)), nn.LeakyReLU(0.1, inplace=True),
            spectral_norm(nn.Conv2d(64, 128, 4, 2, 1)), nn.LeakyReLU(0.1, inplace=True),
            spectral_norm(nn.Conv2d(128, 128, 4, 2, 1)), nn.LeakyReLU(0.1, inplace=True),
            nn.Conv2d(128, 1 + opt.code_dim + opt.n_classes, 4, 1, 0),
        )

    def forward(self, img):
        out = self.main(img).squeeze()
        validity = F.sigmoid(out[:, 0])
        cont = out[:, 1:1 + opt.code_dim]
        cat = F.softmax(out[:, 1 + opt.code_dim:])
        return cat, cont, validity


adversarial_loss = torch.nn.BCELoss()
categorical_loss = torch.nn.CrossEntropyLoss()
continuous_loss = torch.nn.MSELoss()

generator = Generator()
discriminator = Discriminator()
if cuda:
    generator.cuda()
    discriminator.cuda()
    adversarial_loss.cuda()
    categorical_loss.cuda()
    continuous_loss.cuda()

optimizer_G = torch.optim.Adam(generator.parameters(), lr=0.001, betas=(0.5, 0.999))
optimizer_D = torch.optim.Adam(discriminator.parameters(), lr=0.0002, betas=(0.5, 0.999))

FloatTensor = torch.cuda.FloatTensor if cuda else torch.FloatTensor
LongTensor = torch.cuda.LongTensor if cuda else torch.LongTensor

for i in range(opt.n_iter):
    B = opt.batch_size
    real_imgs = torch.tensor(np.random.uniform(-1, 1, (B, 3, 64, 64)), dtype=torch.float32).type(FloatTensor)
    valid = FloatTensor(B).fill_(1.0)
    fake = FloatTensor(B).fill_(0.0)
    z = FloatTensor(np.random.normal(0, 1, (B, opt.latent_dim)))
    sampled = np.random.randint(0, opt.n_classes, B)
    onehot = np.zeros((B, opt.n_classes))
    onehot[range(B), sampled] = 1.0
    labels = FloatTensor(onehot)
    code = FloatTensor(np.random.uniform(-1, 1, (B, opt.code_dim)))

    optimizer_G.zero_grad()
    gen_imgs = generator(z, labels, code)
    pred_label, pred_code, validity = discriminator(gen_imgs)
    g_loss = adversarial_loss(validity, valid) + categorical_loss(pred_label, LongTensor(sampled)) + \
        continuous_loss(pred_code, code)
    g_loss.backward()
    optimizer_G.step()

    optimizer_D.zero_grad()
    _, _, real_pred = discriminator(real_imgs)
    _, _, fake_pred = discriminator(gen_imgs.detach())
    d_loss = (adversarial_loss(real_pred, valid) + adversarial_loss(fake_pred, fake)) / 2
    d_loss.backward()
    optimizer_D.step()

    print(json.dumps({"iter": i, "g_loss": g_loss.item(), "d_loss": d_loss.item(),
                      "G": type(generator.conv_blocks).__module__, "opt": type(optimizer_G).__module__}), flush=True)
