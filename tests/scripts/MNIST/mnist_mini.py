"""A flat training script shaped like the reference's MNIST/EAD-GAN_rpqmnxy.py: generator Linear -> view ->
BatchNorm2d -> Upsample -> Conv3x3 -> BatchNorm2d(C, 0.8) -> LeakyReLU ... -> Tanh, spectral-norm 3x3 stride-2
discriminator / encoder trunks with BatchNorm2d(C, 0.8) and spectral-norm Linear heads, ``nn.Softmax()`` with implicit
dim, LSGAN (MSE) adversarial loss, CrossEntropy on a softmax output, the MLP-approximator affine regulariser from
``from utils_rpqmnxy import *``, ``weights_init_normal`` dispatching on class names, three Adams.  Run UNCHANGED by
tests/test_run_gpu.py with stock PyTorch and under ``python -m eadgan_b200.run``."""
import argparse
import itertools
import json

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Variable
from torch.nn.utils import spectral_norm

from utils_rpqmnxy import *    # noqa: F401,F403

parser = argparse.ArgumentParser()
parser.add_argument("--n_iter", type=int, default=3)
parser.add_argument("--batch_size", type=int, default=16)
parser.add_argument("--latent_dim", type=int, default=62)
parser.add_argument("--code_dim", type=int, default=7)
parser.add_argument("--n_classes", type=int, default=10)
parser.add_argument("--img_size", type=int, default=32)
parser.add_argument("--seed", type=int, default=0)
opt = parser.parse_args()
torch.manual_seed(opt.seed)
np.random.seed(opt.seed)
FloatTensor, LongTensor = torch.cuda.FloatTensor, torch.cuda.LongTensor


def weights_init_normal(m):
    classname = m.__class__.__name__
    if classname.find("Conv") != -1:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find("BatchNorm") != -1:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


class Generator(nn.Module):
    def __init__(self):
        super(Generator, self).__init__()
        self.init_size = opt.img_size // 4
        self.l1 = nn.Sequential(nn.Linear(opt.latent_dim + opt.n_classes + opt.code_dim, 128 * self.init_size ** 2))
        self.conv_blocks = nn.Sequential(
            nn.BatchNorm2d(128), nn.Upsample(scale_factor=2), nn.Conv2d(128, 128, 3, stride=1, padding=1),
            nn.BatchNorm2d(128, 0.8), nn.LeakyReLU(0.2, inplace=True), nn.Upsample(scale_factor=2),
            nn.Conv2d(128, 64, 3, stride=1, padding=1), nn.BatchNorm2d(64, 0.8), nn.LeakyReLU(0.2, inplace=True),
            nn.Conv2d(64, 1, 3, stride=1, padding=1), nn.Tanh())

    def forward(self, noise, labels, code):
        out = self.l1(torch.cat((noise, labels, code), -1))
        return self.conv_blocks(out.view(out.shape[0], 128, self.init_size, self.init_size))


def block(cin, cout, bn):
    layers = [spectral_norm(nn.Conv2d(cin, cout, 3, 2, 1)), nn.LeakyReLU(0.2, inplace=True)]
    return layers + ([nn.BatchNorm2d(cout, 0.8)] if bn else [])


class Discriminator(nn.Module):
    def __init__(self):
        super(Discriminator, self).__init__()
        self.conv_blocks = nn.Sequential(*block(1, 16, False), *block(16, 32, False), *block(32, 64, False), *block(64, 128, False))
        self.adv_layer = nn.Sequential(spectral_norm(nn.Linear(128 * (opt.img_size // 16) ** 2, 1)))

    def forward(self, img):
        out = self.conv_blocks(img)
        return self.adv_layer(out.view(out.shape[0], -1))


class Encoder(nn.Module):
    def __init__(self):
        super(Encoder, self).__init__()
        self.conv_blocks = nn.Sequential(*block(1, 16, False), *block(16, 32, True), *block(32, 64, True), *block(64, 128, True))
        n = 128 * (opt.img_size // 16) ** 2
        self.aux_layer = nn.Sequential(spectral_norm(nn.Linear(n, opt.n_classes)), nn.Softmax())
        self.latent_layer = nn.Sequential(spectral_norm(nn.Linear(n, opt.code_dim)))
        self.noise_layer = nn.Sequential(spectral_norm(nn.Linear(n, opt.latent_dim)))

    def forward(self, img):
        out = self.conv_blocks(img)
        out = out.view(out.shape[0], -1)
        return self.aux_layer(out), self.latent_layer(out), self.noise_layer(out)


class transformation_2D(nn.Module):
    def forward(self, img, matrix_2D):
        grid = F.affine_grid(matrix_2D, img.size())
        return F.grid_sample(img, grid, padding_mode="border")


adversarial_loss, categorical_loss, continuous_loss = torch.nn.MSELoss(), torch.nn.CrossEntropyLoss(), torch.nn.MSELoss()
generator, discriminator, encoder, trans_2D = Generator(), Discriminator(), Encoder(), transformation_2D()
for m in (generator, discriminator, encoder, adversarial_loss, categorical_loss, continuous_loss):
    m.cuda()
generator.apply(weights_init_normal)
discriminator.apply(weights_init_normal)
encoder.apply(weights_init_normal)
optimizer_G = torch.optim.Adam(generator.parameters(), lr=0.0002, betas=(0.5, 0.999))
optimizer_D = torch.optim.Adam(discriminator.parameters(), lr=0.0002, betas=(0.5, 0.999))
optimizer_info = torch.optim.Adam(itertools.chain(generator.parameters(), encoder.parameters()), lr=0.0002, betas=(0.5, 0.999))

for it in range(opt.n_iter):
    B = opt.batch_size
    real_imgs = Variable(FloatTensor(np.random.uniform(-1, 1, (B, 1, opt.img_size, opt.img_size))))
    valid = Variable(FloatTensor(B, 1).fill_(1.0), requires_grad=False)
    fake = Variable(FloatTensor(B, 1).fill_(0.0), requires_grad=False)
    z = Variable(FloatTensor(np.random.normal(0, 1, (B, opt.latent_dim))))
    sampled = np.random.randint(0, opt.n_classes, B)
    onehot = np.zeros((B, opt.n_classes))
    onehot[range(B), sampled] = 1.0
    label_input = Variable(FloatTensor(onehot))
    code_input = Variable(FloatTensor(np.random.uniform(-1, 1, (B, opt.code_dim))))
    scaled_img = trans_2D(real_imgs, get_matrix(code_input)[:, 0:2])

    optimizer_G.zero_grad()
    gen_imgs = generator(z, label_input, code_input)
    g_loss = adversarial_loss(discriminator(gen_imgs), valid)
    g_loss.backward()
    optimizer_G.step()

    optimizer_D.zero_grad()
    d_loss = (adversarial_loss(discriminator(scaled_img), valid) + adversarial_loss(discriminator(gen_imgs.detach()), fake)) / 2
    d_loss.backward()
    optimizer_D.step()

    optimizer_info.zero_grad()
    gen_imgs = generator(z, label_input, code_input)
    pred_label, pred_code, _ = encoder(gen_imgs)
    info_loss = categorical_loss(pred_label, Variable(LongTensor(sampled))) + 0.1 * continuous_loss(pred_code, code_input)
    _, transform_code, _ = encoder(scaled_img)
    _, real_code, _ = encoder(real_imgs)
    info_loss = info_loss + 0.1 * continuous_loss(affine_regularizer(real_code, transform_code), code_input)
    info_loss.backward()
    optimizer_info.step()

    print(json.dumps({"iter": it, "g_loss": g_loss.item(), "d_loss": d_loss.item(), "info_loss": info_loss.item(),
                      "G": type(generator.conv_blocks).__module__, "opt": type(optimizer_G).__module__,
                      "utils": get_matrix.__module__}), flush=True)
